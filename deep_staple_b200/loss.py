"""Data-parameter-weighted segmentation loss on B200 -- call-site replacement for the inline loss
of the reference's ``train_DL`` (multimodallearning/deep_staple, main_deep_staple.py):

* ``weighted_ce``        == ``nn.CrossEntropyLoss(class_weights)(logits, b_seg_modified)``      :716 (+ backward :719)
* ``dp_weighted_loss``   == lines :738-759 (per-instance CE, sigmoid/mean-normalised data-parameter
                            weight, fixed weighting, risk regularisation) with the backward of :762
                            into the data parameters (and into ``dp_logits`` when it requires grad)

Both run as ONE fused streaming kernel launch (``ds_dp_loss_fwd_bwd`` / ``ds_weighted_ce_fwd_bwd``,
include/deep_staple_b200.h): forward statistics and the logits gradient are produced in the same pass
over HBM, because the instance weights depend only on the data parameters and are known beforehand.
The autograd ``backward`` therefore only hands out what the forward pass already computed.

No CPU fallback: tensors must live on a B200.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple, Union

import torch

from . import _binding as _b

__all__ = ["dp_weighted_loss", "weighted_ce", "DPLossOutput", "loss_statistics", "backward_logits"]

_workspaces = {}


def _workspace(device: torch.device, B: int, C: int) -> torch.Tensor:
    """Zero-initialised scratch (ticket counters + per-CTA partials), one per (device, stream, B, C).
    The kernels leave the counters zero again, so it is cleared exactly once."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, B, C)
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = int(_b.load().ds_dp_loss_workspace_bytes(B, C))
        ws = torch.zeros((nbytes + 15) // 16 * 2, dtype=torch.int64, device=device)
        _workspaces[key] = ws
    return ws


def _scale_tensor(grad_scaler, device) -> Optional[torch.Tensor]:
    """Device scalar the fused pass multiplies into the gradients: the GradScaler's current scale
    (``scaler_dp.scale(dp_loss).backward()``, main_deep_staple.py:762).  ``None`` means 1."""
    if grad_scaler is None:
        return None
    if isinstance(grad_scaler, torch.Tensor):
        return grad_scaler.detach().to(device=device, dtype=torch.float32).reshape(())
    if isinstance(grad_scaler, (int, float)):
        return torch.tensor(float(grad_scaler), dtype=torch.float32, device=device)
    if hasattr(grad_scaler, "is_enabled") and not grad_scaler.is_enabled():
        return None
    scale = getattr(grad_scaler, "_scale", None)
    if scale is None:  # lazily created by the first scale() call
        return torch.tensor(float(grad_scaler._init_scale), dtype=torch.float32, device=device)
    return scale.detach().to(device=device, dtype=torch.float32).reshape(())


def _flat_inputs(logits: torch.Tensor, target: torch.Tensor):
    _b.require_cuda(logits, "logits")
    _b.require_cuda(target, "target")
    if logits.dim() < 3:
        raise ValueError(f"logits must be [B, C, *spatial], got {tuple(logits.shape)}")
    B, C = logits.shape[0], logits.shape[1]
    if tuple(target.shape) != (B, *logits.shape[2:]):
        raise ValueError(f"target shape {tuple(target.shape)} does not match logits {tuple(logits.shape)}")
    if target.dtype not in (torch.int64, torch.uint8):
        raise ValueError(f"target must be int64 or uint8, got {target.dtype}")
    V = 1
    for s in logits.shape[2:]:
        V *= s
    return logits.contiguous(), target.contiguous(), B, C, V


@dataclass
class DPLossOutput:
    """Everything the fused pass produced (all tensors on the device, no host sync)."""

    dp_loss: torch.Tensor  # 0-dim fp32, autograd-connected                      (main_deep_staple.py:757/759)
    ce_loss: torch.Tensor  # 0-dim fp32 class-weighted CE value (0 without class_weights)   (:716)
    per_instance_ce: torch.Tensor  # [B]  l_b                                                (:739)
    fg_pred: torch.Tensor  # [B]  n_b = #voxels predicted foreground                         (:751)
    instance_weight: torch.Tensor  # [B]  w_b                                                (:743-748)
    inst_stats: torch.Tensor  # [B, C, 4] float64: sum-nll / label / pred / TP counts (hard-Dice statistics)
    d_bare_weight: Optional[torch.Tensor] = None  # [B] d(scale * dp_loss)/d(bare_weight), for optim.dp_optimizer_step


class _DPLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dp_logits, bare_weight, target, fixed_b, use_risk, scale, class_w, dice_target):
        z, t, B, C, V = _flat_inputs(dp_logits, target)
        dev = z.device
        bare = bare_weight.detach().to(torch.float32).reshape(-1).contiguous()
        if bare.numel() != B:
            raise ValueError(f"{B} instances in the batch but {bare.numel()} data parameters")
        fixed = None if fixed_b is None else fixed_b.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if fixed is not None and fixed.numel() != B:
            raise ValueError(f"{B} instances in the batch but {fixed.numel()} fixed weights")
        cw = None if class_w is None else class_w.detach().to(device=dev, dtype=torch.float32).contiguous()
        dt = None
        if dice_target is not None:
            dt = dice_target.contiguous()
            if dt.dtype != t.dtype or dt.shape != t.shape:
                raise ValueError("dice_target must match target in dtype and shape")
        need_dlogits = ctx.needs_input_grad[0]
        dlogits = torch.empty_like(z) if need_dlogits else None
        inst_stats = torch.empty((B, C, _b.DS_STATS), dtype=torch.float64, device=dev)
        loss_out = torch.empty(2, dtype=torch.float32, device=dev)
        d_bare = torch.empty(B, dtype=torch.float32, device=dev)
        per_inst = torch.empty((B, _b.DS_INST), dtype=torch.float32, device=dev)
        lib = _b.load()
        with torch.cuda.device(dev):
            ws = _workspace(dev, B, C)
            rc = lib.ds_dp_loss_fwd_bwd(
                _b.ptr(z), _b.logit_dtype(z), _b.ptr(t), _b.label_dtype(t), _b.ptr(dt), B, C, V,
                _b.ptr(bare), _b.ptr(fixed), B, 0, _b.ptr(cw), 1 if use_risk else 0, _b.ptr(scale),
                _b.ptr(dlogits), _b.ptr(inst_stats), _b.ptr(loss_out), _b.ptr(d_bare), _b.ptr(per_inst),
                _b.ptr(ws), ws.numel() * 8, _b.stream_ptr(dev),
            )
        _b.check(rc, "ds_dp_loss_fwd_bwd")
        ctx.bare_shape = bare_weight.shape
        ctx.logits_shape = dp_logits.shape
        ctx.assumed = scale
        ctx.save_for_backward(d_bare, dlogits)
        ce, pi = loss_out[1], per_inst
        ctx.mark_non_differentiable(ce, pi, inst_stats, d_bare)
        return loss_out[0], ce, pi, inst_stats, d_bare

    @staticmethod
    def backward(ctx, grad_loss, _g_ce, _g_pi, _g_stats, _g_d):
        d_bare, dlogits = ctx.saved_tensors
        assumed = ctx.assumed
        g = grad_loss.detach().to(torch.float32).reshape(())
        # d_bare / dlogits already carry the scale assumed in the fused pass; reconcile with the
        # gradient that actually arrived (equal in the GradScaler flow -> the kernel reads nothing)
        ratio = g if assumed is None else g / assumed
        grad_bare = (d_bare * ratio).reshape(ctx.bare_shape)
        grad_logits = None
        if dlogits is not None and ctx.needs_input_grad[0]:
            one = assumed if assumed is not None else torch.ones((), dtype=torch.float32, device=g.device)
            lib = _b.load()
            with torch.cuda.device(dlogits.device):
                rc = lib.ds_rescale_inplace(_b.ptr(dlogits), _b.logit_dtype(dlogits), dlogits.numel(), _b.ptr(g),
                                            _b.ptr(one), _b.stream_ptr(dlogits.device))
            _b.check(rc, "ds_rescale_inplace")
            grad_logits = dlogits.view(ctx.logits_shape)
        return grad_logits, grad_bare, None, None, None, None, None, None


def dp_weighted_loss(
    dp_logits: torch.Tensor,
    b_seg_modified: torch.Tensor,
    bare_weight: Union[torch.Tensor, Tuple[torch.nn.Embedding, torch.Tensor]],
    fixed_weighting: Optional[torch.Tensor] = None,
    use_risk_regularization: bool = True,
    *,
    grad_scaler=None,
    class_weights: Optional[torch.Tensor] = None,
    dice_target: Optional[torch.Tensor] = None,
    return_all: bool = False,
):
    """Replacement for main_deep_staple.py:738-759 (forward) + :762 (backward).

    dp_logits        ``[B, C, *spatial]`` fp32 / fp16 / bf16 logits.  If it requires grad (in-line
                     mode, :731-736) its gradient is produced by the same pass; in the out-of-line
                     mode (:724-729, ``requires_grad == False``) only statistics are computed.
    b_seg_modified   ``[B, *spatial]`` int64 (reference) or uint8 labels.
    bare_weight      ``embedding(b_idxs_dataset).squeeze()`` (:741) -- pass the tensor the reference
                     computes (its autograd graph then delivers the sparse gradient to
                     ``embedding.weight``), or the pair ``(embedding, b_idxs_dataset)``.
    fixed_weighting  ``fixed_weighting[b_idxs_dataset]`` (:748) or ``None`` (``use_fixed_weighting`` off).
    grad_scaler      the ``GradScaler`` (or scale tensor / float) that will scale this loss (:762);
                     lets the fused pass emit already-scaled gradients.  Optional: a different
                     incoming gradient is reconciled in ``backward`` at the cost of one more pass.
    class_weights    optional ``[C]``: also return the class-weighted CE value of :716 (``return_all``).
    dice_target      optional clean label map (``b_seg``, :785) for the true-positive statistic.
    returns          0-dim fp32 ``dp_loss`` -- or a ``DPLossOutput`` when ``return_all``.
    """
    if isinstance(bare_weight, (tuple, list)):
        embedding, idx = bare_weight
        bare_weight = embedding(idx).squeeze()  # :741 (sparse=True embedding -> sparse grad)
    scale = _scale_tensor(grad_scaler, dp_logits.device)
    loss, ce, per_inst, stats, d_bare = _DPLoss.apply(
        dp_logits, bare_weight, b_seg_modified, fixed_weighting, bool(use_risk_regularization), scale,
        class_weights, dice_target,
    )
    if not return_all:
        return loss
    return DPLossOutput(
        dp_loss=loss, ce_loss=ce, per_instance_ce=per_inst[:, _b.DS_INST_CE], fg_pred=per_inst[:, _b.DS_INST_FG],
        instance_weight=per_inst[:, _b.DS_INST_WEIGHT], inst_stats=stats, d_bare_weight=d_bare,
    )


def loss_statistics(logits: torch.Tensor, target: torch.Tensor, dice_target: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Forward statistics only: ``[B, C, 4]`` float64 (sum-nll, label, pred, TP counts)."""
    z, t, B, C, V = _flat_inputs(logits, target)
    zero = torch.zeros(B, dtype=torch.float32, device=z.device)
    out = dp_weighted_loss(z.detach(), t, zero, None, False, dice_target=dice_target, return_all=True)
    return out.inst_stats


def label_statistics(pred: torch.Tensor, target: Optional[torch.Tensor], num_classes: int) -> torch.Tensor:
    """``[B, C, 4]`` float64 counts from label maps only (ds_label_stats): LABEL = #(target==c),
    PRED = #(pred==c), TP = #(pred==c and target==c).  Replaces the counting of
    main_deep_staple.py:646-648 and deep_staple/metrics.py:93-101."""
    _b.require_cuda(pred, "pred")
    p = pred.contiguous()
    if p.dtype not in (torch.int64, torch.uint8):
        p = p.long()
    t = None
    if target is not None:
        _b.require_cuda(target, "target")
        t = target.contiguous().to(p.dtype)
        if t.shape != p.shape:
            raise ValueError("pred and target must have the same shape")
    B = p.shape[0]
    V = p[0].numel()
    stats = torch.empty((B, num_classes, _b.DS_STATS), dtype=torch.float64, device=p.device)
    lib = _b.load()
    with torch.cuda.device(p.device):
        ws = _workspace(p.device, B, num_classes)
        rc = lib.ds_label_stats(_b.ptr(p), _b.label_dtype(p), _b.ptr(t), _b.label_dtype(p), B, num_classes, V,
                                _b.ptr(stats), _b.ptr(ws), ws.numel() * 8, _b.stream_ptr(p.device))
    _b.check(rc, "ds_label_stats")
    return stats


class _WeightedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, class_w, scale, class_counts):
        z, t, B, C, V = _flat_inputs(logits, target)
        dev = z.device
        cw = class_w.detach().to(device=dev, dtype=torch.float32).contiguous()
        if cw.numel() != C:
            raise ValueError(f"{C} classes but {cw.numel()} class weights")
        need_dlogits = ctx.needs_input_grad[0]
        counts = None
        if need_dlogits:
            # the normaliser sum_j cw[t_j] must be known before the fused pass: label histogram first
            if class_counts is None:
                class_counts = label_statistics(t, None, C)[:, :, _b.DS_STAT_PRED].sum(0)
            counts = class_counts.detach().to(device=dev, dtype=torch.float64).contiguous()
        dlogits = torch.empty_like(z) if need_dlogits else None
        inst_stats = torch.empty((B, C, _b.DS_STATS), dtype=torch.float64, device=dev)
        loss_out = torch.empty(2, dtype=torch.float32, device=dev)
        lib = _b.load()
        with torch.cuda.device(dev):
            ws = _workspace(dev, B, C)
            rc = lib.ds_weighted_ce_fwd_bwd(
                _b.ptr(z), _b.logit_dtype(z), _b.ptr(t), _b.label_dtype(t), B, C, V, _b.ptr(cw), _b.ptr(counts),
                _b.ptr(scale), _b.ptr(dlogits), _b.ptr(inst_stats), _b.ptr(loss_out), _b.ptr(ws), ws.numel() * 8,
                _b.stream_ptr(dev),
            )
        _b.check(rc, "ds_weighted_ce_fwd_bwd")
        ctx.assumed = scale
        ctx.logits_shape = logits.shape
        ctx.save_for_backward(dlogits)
        ctx.mark_non_differentiable(inst_stats)
        return loss_out[1], inst_stats

    @staticmethod
    def backward(ctx, grad_loss, _g_stats):
        (dlogits,) = ctx.saved_tensors
        if dlogits is None:
            return None, None, None, None, None
        g = grad_loss.detach().to(torch.float32).reshape(())
        one = ctx.assumed if ctx.assumed is not None else torch.ones((), dtype=torch.float32, device=g.device)
        lib = _b.load()
        with torch.cuda.device(dlogits.device):
            rc = lib.ds_rescale_inplace(_b.ptr(dlogits), _b.logit_dtype(dlogits), dlogits.numel(), _b.ptr(g), _b.ptr(one),
                                        _b.stream_ptr(dlogits.device))
        _b.check(rc, "ds_rescale_inplace")
        return dlogits.view(ctx.logits_shape), None, None, None, None


def weighted_ce(logits: torch.Tensor, target: torch.Tensor, class_weights: torch.Tensor, *, grad_scaler=None,
                class_counts: Optional[torch.Tensor] = None, return_stats: bool = False):
    """``nn.CrossEntropyLoss(class_weights)(logits, target)`` (main_deep_staple.py:716) with its
    backward (:719) fused into the same pass.  ``class_counts`` (``[C]`` label histogram of the batch)
    may be supplied to skip the label pre-pass (e.g. the all-reduced global histogram under data
    parallelism)."""
    scale = _scale_tensor(grad_scaler, logits.device)
    loss, stats = _WeightedCE.apply(logits, target, class_weights, scale, class_counts)
    return (loss, stats) if return_stats else loss


def backward_logits(logits: torch.Tensor, target: torch.Tensor, instance_weight: torch.Tensor,
                    grad_out: Union[torch.Tensor, float] = 1.0) -> torch.Tensor:
    """Two-pass backward: ``grad_out * w_b / V * (softmax(z) - onehot(t))`` recomputed from the logits
    (ds_dp_loss_backward_logits)."""
    z, t, B, C, V = _flat_inputs(logits, target)
    w = instance_weight.detach().to(device=z.device, dtype=torch.float32).reshape(-1).contiguous()
    g = _scale_tensor(grad_out, z.device)
    out = torch.empty_like(z)
    lib = _b.load()
    with torch.cuda.device(z.device):
        rc = lib.ds_dp_loss_backward_logits(_b.ptr(z), _b.logit_dtype(z), _b.ptr(t), _b.label_dtype(t), B, C, V,
                                            _b.ptr(w), 1, _b.ptr(g), _b.ptr(out), _b.stream_ptr(z.device))
    _b.check(rc, "ds_dp_loss_backward_logits")
    return out.view(logits.shape)
