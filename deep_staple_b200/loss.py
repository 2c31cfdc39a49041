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

from typing import Optional, Tuple, Union

import torch

from . import _binding as _b

__all__ = ["dp_weighted_loss", "weighted_ce", "DPLossOutput", "loss_statistics", "backward_logits"]

_workspaces = {}
_F32, _F64 = torch.float32, torch.float64


def _workspace(device: torch.device, B: int, C: int) -> torch.Tensor:
    """Zero-initialised scratch (ticket counters + exact accumulators), one per (device, stream, B, C).
    The kernels leave it all-zero again, so it is cleared exactly once.  The cache is bounded: a training run
    uses a handful of (stream, batch shape) combinations; beyond 64 entries the oldest is dropped."""
    key = (device.index, _b.stream_ptr(device), B, C)
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = int(_b.load().ds_dp_loss_workspace_bytes(B, C))
        ws = torch.zeros((nbytes + 15) // 16 * 2, dtype=torch.int64, device=device)
        if len(_workspaces) >= 64:
            _workspaces.pop(next(iter(_workspaces)))
        _workspaces[key] = ws
    return ws


_scale_constants = {}


def _scale_tensor(grad_scaler, device) -> Optional[torch.Tensor]:
    """Device scalar the fused pass multiplies into the gradients: the GradScaler's current scale
    (``scaler_dp.scale(dp_loss).backward()``, main_deep_staple.py:762).  ``None`` means 1."""
    if grad_scaler is None:
        return None
    if isinstance(grad_scaler, torch.Tensor):
        t = grad_scaler
        if t.dtype is _F32 and t.dim() == 0 and t.device == device and not t.requires_grad:
            return t
        return t.detach().to(device=device, dtype=_F32).reshape(())
    if isinstance(grad_scaler, (int, float)):
        return _scale_constant(float(grad_scaler), device)
    if hasattr(grad_scaler, "is_enabled") and not grad_scaler.is_enabled():
        return None
    scale = getattr(grad_scaler, "_scale", None)
    if scale is None:  # lazily created by the first scale() call
        return _scale_constant(float(grad_scaler._init_scale), device)
    if scale.dtype is _F32 and scale.dim() == 0 and scale.device == device:
        return scale
    return scale.detach().to(device=device, dtype=_F32).reshape(())


def _scale_constant(value: float, device) -> torch.Tensor:
    """A python-number scale as a cached device scalar (building one per call is a blocking H2D copy)."""
    key = (value, device.index)
    t = _scale_constants.get(key)
    if t is None:
        if len(_scale_constants) >= 64:
            _scale_constants.clear()
        t = _scale_constants[key] = torch.tensor(value, dtype=_F32, device=device)
    return t


def _flat_inputs(logits: torch.Tensor, target: torch.Tensor):
    _b.require_cuda(logits, "logits")
    _b.require_cuda(target, "target")
    shape = logits.shape
    if len(shape) < 3:
        raise ValueError(f"logits must be [B, C, *spatial], got {tuple(shape)}")
    B, C = shape[0], shape[1]
    if target.shape[1:] != shape[2:] or target.shape[0] != B:
        raise ValueError(f"target shape {tuple(target.shape)} does not match logits {tuple(shape)}")
    if target.dtype not in (torch.int64, torch.uint8):
        raise ValueError(f"target must be int64 or uint8, got {target.dtype}")
    V = 1
    for s in shape[2:]:
        V *= s
    return (logits if logits.is_contiguous() else logits.contiguous(),
            target if target.is_contiguous() else target.contiguous(), B, C, V)


def _f32_vector(t: Optional[torch.Tensor], device, n: int, what: str) -> Optional[torch.Tensor]:
    """``t`` as a contiguous fp32 ``[n]`` device vector -- without launching anything when it already is one."""
    if t is None:
        return None
    if t.requires_grad:
        t = t.detach()
    if t.dtype is not _F32 or t.device != device:
        t = t.to(device=device, dtype=_F32)
    if t.dim() != 1:
        t = t.reshape(-1)
    if not t.is_contiguous():
        t = t.contiguous()
    if t.shape[0] != n:
        raise ValueError(f"{n} expected but {t.shape[0]} {what}")
    return t


_layouts = {}


def _arena_layout(B: int, C: int):
    """One allocation per call holds every small output of the fused pass, as fp32 words:
    ``[B, C, 4]`` float64 statistics | dp_loss, ce_loss | d_bare_weight ``[B]`` | per-instance ``[B, 4]``."""
    lay = _layouts.get((B, C))
    if lay is None:
        n_stats = 2 * B * C * _b.DS_STATS
        o_loss = n_stats
        o_dbare = o_loss + 2
        o_inst = o_dbare + B + (B & 1)  # keep the rows 8-byte aligned
        lay = _layouts[(B, C)] = (n_stats, o_loss, o_dbare, o_inst, o_inst + B * _b.DS_INST)
    return lay


class DPLossOutput:
    """Everything the fused pass produced (all tensors on the device, no host sync).  The fields are views of ONE
    arena allocation, created on first access.

    dp_loss          0-dim fp32, autograd-connected                                  (main_deep_staple.py:757/759)
    ce_loss          0-dim fp32 class-weighted CE value (0 without class_weights)     (:716)
    per_instance_ce  [B]  l_b                                                         (:739)
    fg_pred          [B]  n_b = #voxels predicted foreground                          (:751)
    instance_weight  [B]  w_b                                                         (:743-748)
    inst_stats       [B, C, 4] float64: sum-nll / label / pred / TP counts (hard-Dice statistics)
    d_bare_weight    [B] d(scale * dp_loss)/d(bare_weight), for optim.dp_optimizer_step
    """

    _FIELDS = ("ce_loss", "per_instance_ce", "fg_pred", "instance_weight", "inst_stats", "d_bare_weight")

    def __init__(self, dp_loss, arena=None, B=None, C=None, **fields):
        self.dp_loss = dp_loss
        self._arena, self._B, self._C = arena, B, C
        for k, v in fields.items():
            if k not in self._FIELDS:
                raise TypeError(f"unknown field {k}")
            setattr(self, k, v)

    def __getattr__(self, name):  # only reached for fields not materialised yet
        arena = self.__dict__.get("_arena")
        if arena is None or name not in DPLossOutput._FIELDS:
            if name == "d_bare_weight":
                return None
            raise AttributeError(name)
        B, C = self._B, self._C
        n_stats, o_loss, o_dbare, o_inst, _ = _arena_layout(B, C)
        if name == "ce_loss":
            v = arena[o_loss + 1]
        elif name == "inst_stats":
            v = arena[:n_stats].view(_F64).view(B, C, _b.DS_STATS)
        elif name == "d_bare_weight":
            v = arena[o_dbare:o_dbare + B]
        else:
            col = {"per_instance_ce": _b.DS_INST_CE, "fg_pred": _b.DS_INST_FG, "instance_weight": _b.DS_INST_WEIGHT}[name]
            v = arena[o_inst:o_inst + B * _b.DS_INST].view(B, _b.DS_INST)[:, col]
        setattr(self, name, v)
        return v


def _launch_dp_loss(z, t, B, C, V, bare, fixed, cw, dt, use_risk, scale, dlogits):
    """One ``ds_dp_loss_fwd_bwd`` launch on the caller's stream; returns the output arena."""
    dev = z.device
    arena = torch.empty(_arena_layout(B, C)[4], dtype=_F32, device=dev)
    n_stats, o_loss, o_dbare, o_inst, _ = _arena_layout(B, C)
    base = arena.data_ptr()
    ws = _workspace(dev, B, C)
    lib = _b.load()
    args = (z.data_ptr(), _b.logit_dtype(z), t.data_ptr(), _b.label_dtype(t), _b.ptr(dt), B, C, V,
            bare.data_ptr(), _b.ptr(fixed), B, 0, _b.ptr(cw), 1 if use_risk else 0, _b.ptr(scale), _b.ptr(dlogits),
            base, base + 4 * o_loss, base + 4 * o_dbare, base + 4 * o_inst, ws.data_ptr(), ws.numel() * 8,
            _b.stream_ptr(dev))
    if torch.cuda.current_device() == dev.index:
        rc = lib.ds_dp_loss_fwd_bwd(*args)
    else:
        with torch.cuda.device(dev):
            rc = lib.ds_dp_loss_fwd_bwd(*args)
    if rc:
        _b.check(rc, "ds_dp_loss_fwd_bwd")
    return arena


def _prepare_dp(dp_logits, bare_weight, target, fixed_b, class_w, dice_target):
    z, t, B, C, V = _flat_inputs(dp_logits, target)
    dev = z.device
    bare = _f32_vector(bare_weight, dev, B, "data parameters for that many instances")
    fixed = _f32_vector(fixed_b, dev, B, "fixed weights for that many instances")
    cw = _f32_vector(class_w, dev, C, "class weights for that many classes")
    dt = None
    if dice_target is not None:
        dt = dice_target if dice_target.is_contiguous() else dice_target.contiguous()
        if dt.dtype != t.dtype or dt.shape != t.shape:
            raise ValueError("dice_target must match target in dtype and shape")
    return z, t, B, C, V, bare, fixed, cw, dt


class _DPLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dp_logits, bare_weight, target, fixed_b, use_risk, scale, class_w, dice_target, holder):
        z, t, B, C, V, bare, fixed, cw, dt = _prepare_dp(dp_logits, bare_weight, target, fixed_b, class_w, dice_target)
        need_dlogits = ctx.needs_input_grad[0]
        # The logits gradient is written by the same pass, already multiplied by the scale backward will see.
        # Half-precision logits WITHOUT that hint cannot take this route: unscaled gradients (w_b / V ~ 5e-7 for
        # 128^3) would be stored as fp16 subnormals before the scale multiplies them -- they are recomputed from
        # the logits in backward instead (ds_dp_loss_backward_logits), with the gradient that really arrived.
        fused_grad = need_dlogits and (scale is not None or z.dtype is _F32)
        dlogits = torch.empty_like(z) if fused_grad else None
        arena = _launch_dp_loss(z, t, B, C, V, bare, fixed, cw, dt, use_risk, scale, dlogits)
        holder.append(arena)
        ctx.bare_shape = bare_weight.shape
        ctx.logits_shape = dp_logits.shape
        ctx.assumed = scale          # scale inside d_bare_weight (never rewritten)
        ctx.dlogits_scale = scale    # scale inside the stored dlogits (rewritten by a reconciling backward)
        ctx.dims = (B, C, V)
        ctx.two_pass = need_dlogits and not fused_grad
        if ctx.two_pass:
            ctx.save_for_backward(arena, z, t)
        else:
            ctx.save_for_backward(arena, dlogits)
        return arena[_arena_layout(B, C)[1]]

    @staticmethod
    def backward(ctx, grad_loss):
        B, C, V = ctx.dims
        saved = ctx.saved_tensors
        arena = saved[0]
        dev = arena.device
        g = grad_loss
        if g.dtype is not _F32 or g.dim() != 0 or g.requires_grad:
            g = g.detach().to(_F32).reshape(())
        _, _, o_dbare, o_inst, _ = _arena_layout(B, C)
        grad_bare = torch.empty(B, dtype=_F32, device=dev)
        dlogits = None if ctx.two_pass else saved[1]
        want_dlogits = ctx.needs_input_grad[0] and (dlogits is not None or ctx.two_pass)
        lib = _b.load()
        base = arena.data_ptr()
        fused = dlogits if want_dlogits else None
        with torch.cuda.device(dev):
            # one launch: d_bare_weight x (g / assumed) and -- only if g differs from the scale already inside the
            # stored dlogits -- their rescale (in place; the scale they now carry is remembered for a second backward)
            if fused is not None and ctx.dlogits_scale is not ctx.assumed:
                rc = lib.ds_rescale_inplace(fused.data_ptr(), _b.logit_dtype(fused), fused.numel(), g.data_ptr(),
                                            _b.ptr(ctx.dlogits_scale) or _scale_constant(1.0, dev).data_ptr(),
                                            _b.stream_ptr(dev))
                _b.check(rc, "ds_rescale_inplace")
                fused_arg = None
            else:
                fused_arg = fused
            rc = lib.ds_dp_loss_backward_reconcile(
                _b.ptr(fused_arg), _b.logit_dtype(fused_arg) if fused_arg is not None else _b.DS_F32,
                fused_arg.numel() if fused_arg is not None else 0, base + 4 * o_dbare, grad_bare.data_ptr(), B,
                g.data_ptr(), _b.ptr(ctx.assumed), _b.stream_ptr(dev))
            if rc:
                _b.check(rc, "ds_dp_loss_backward_reconcile")
            grad_logits = None
            if fused is not None:
                ctx.dlogits_scale = g
                grad_logits = fused.view(ctx.logits_shape)
            elif want_dlogits:  # two-pass route (half logits without a scale hint)
                z, t = saved[1], saved[2]
                grad_logits = torch.empty_like(z)
                rc = lib.ds_dp_loss_backward_logits(
                    z.data_ptr(), _b.logit_dtype(z), t.data_ptr(), _b.label_dtype(t), B, C, V,
                    base + 4 * (o_inst + _b.DS_INST_WEIGHT), _b.DS_INST, g.data_ptr(), grad_logits.data_ptr(),
                    _b.stream_ptr(dev))
                _b.check(rc, "ds_dp_loss_backward_logits")
                grad_logits = grad_logits.view(ctx.logits_shape)
        if grad_bare.shape != ctx.bare_shape:
            grad_bare = grad_bare.reshape(ctx.bare_shape)
        return grad_logits, grad_bare, None, None, None, None, None, None, None


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
                     lets the fused pass emit already-scaled gradients.  Without it fp32 logits gradients are
                     reconciled in ``backward`` (one more pass over them); half-precision logits gradients are
                     then recomputed from the logits in ``backward`` -- pass it whenever AMP is on.
    class_weights    optional ``[C]``: also return the class-weighted CE value of :716 (``return_all``).
    dice_target      optional clean label map (``b_seg``, :785) for the true-positive statistic.
    returns          0-dim fp32 ``dp_loss`` -- or a ``DPLossOutput`` when ``return_all``.
    """
    if isinstance(bare_weight, (tuple, list)):
        embedding, idx = bare_weight
        bare_weight = embedding(idx).squeeze()  # :741 (sparse=True embedding -> sparse grad)
    scale = _scale_tensor(grad_scaler, dp_logits.device)
    if torch.is_grad_enabled() and (dp_logits.requires_grad or bare_weight.requires_grad):
        holder = []
        loss = _DPLoss.apply(dp_logits, bare_weight, b_seg_modified, fixed_weighting, bool(use_risk_regularization), scale,
                             class_weights, dice_target, holder)
        arena = holder[0]
    else:  # nothing to differentiate: no autograd node at all
        z, t, B, C, V, bare, fixed, cw, dt = _prepare_dp(dp_logits, bare_weight, b_seg_modified, fixed_weighting,
                                                         class_weights, dice_target)
        arena = _launch_dp_loss(z, t, B, C, V, bare, fixed, cw, dt, bool(use_risk_regularization), scale, None)
        loss = arena[_arena_layout(B, C)[1]]
    if not return_all:
        return loss
    return DPLossOutput(loss, arena, dp_logits.shape[0], dp_logits.shape[1])


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


def _launch_weighted_ce(z, t, B, C, V, cw, counts, scale, dlogits):
    """One ``ds_weighted_ce_fwd_bwd`` launch; returns the arena (statistics | dp_loss slot, ce_loss | ...)."""
    dev = z.device
    n_stats, o_loss, _, _, total = _arena_layout(B, C)
    arena = torch.empty(total, dtype=_F32, device=dev)
    base = arena.data_ptr()
    ws = _workspace(dev, B, C)
    lib = _b.load()
    args = (z.data_ptr(), _b.logit_dtype(z), t.data_ptr(), _b.label_dtype(t), B, C, V, cw.data_ptr(), _b.ptr(counts),
            _b.ptr(scale), _b.ptr(dlogits), base, base + 4 * o_loss, ws.data_ptr(), ws.numel() * 8, _b.stream_ptr(dev))
    if torch.cuda.current_device() == dev.index:
        rc = lib.ds_weighted_ce_fwd_bwd(*args)
    else:
        with torch.cuda.device(dev):
            rc = lib.ds_weighted_ce_fwd_bwd(*args)
    if rc:
        _b.check(rc, "ds_weighted_ce_fwd_bwd")
    return arena


def _ce_counts(class_counts, t, C, dev):
    # the normaliser sum_j cw[t_j] must be known before the fused pass: label histogram first (one more read of the
    # labels) unless the caller has it (e.g. from the data loader, or all-reduced under data parallelism)
    if class_counts is None:
        return label_statistics(t, None, C)[:, :, _b.DS_STAT_PRED].sum(0)
    c = class_counts.detach() if class_counts.requires_grad else class_counts
    if c.dtype is not _F64 or c.device != dev:
        c = c.to(device=dev, dtype=_F64)
    return c if c.is_contiguous() else c.contiguous()


class _WeightedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, class_w, scale, class_counts, holder):
        z, t, B, C, V = _flat_inputs(logits, target)
        dev = z.device
        cw = _f32_vector(class_w, dev, C, "class weights for that many classes")
        # half-precision logits without a scale hint: gradients of magnitude cw / sum(cw) ~ 1e-7 would be stored as
        # fp16 subnormals -- the fused pass is then run in backward, with the gradient that really arrived
        deferred = scale is None and z.dtype is not _F32
        dlogits = None if deferred else torch.empty_like(z)
        counts = _ce_counts(class_counts, t, C, dev)
        arena = _launch_weighted_ce(z, t, B, C, V, cw, counts, scale, dlogits)
        holder.append(arena)
        ctx.dlogits_scale = scale
        ctx.logits_shape = logits.shape
        ctx.dims = (B, C, V)
        if deferred:
            ctx.save_for_backward(z, t, cw, counts)
        else:
            ctx.save_for_backward(dlogits)
        ctx.deferred = deferred
        return arena[_arena_layout(B, C)[1] + 1]

    @staticmethod
    def backward(ctx, grad_loss):
        g = grad_loss
        if g.dtype is not _F32 or g.dim() != 0 or g.requires_grad:
            g = g.detach().to(_F32).reshape(())
        if ctx.deferred:
            z, t, cw, counts = ctx.saved_tensors
            B, C, V = ctx.dims
            dlogits = torch.empty_like(z)
            _launch_weighted_ce(z, t, B, C, V, cw, counts, g, dlogits)
            return dlogits.view(ctx.logits_shape), None, None, None, None, None
        (dlogits,) = ctx.saved_tensors
        dev = dlogits.device
        one = ctx.dlogits_scale if ctx.dlogits_scale is not None else _scale_constant(1.0, dev)
        lib = _b.load()
        args = (dlogits.data_ptr(), _b.logit_dtype(dlogits), dlogits.numel(), g.data_ptr(), one.data_ptr(), _b.stream_ptr(dev))
        if torch.cuda.current_device() == dev.index:
            rc = lib.ds_rescale_inplace(*args)  # reads two scalars and returns when they are equal (the GradScaler flow)
        else:
            with torch.cuda.device(dev):
                rc = lib.ds_rescale_inplace(*args)
        if rc:
            _b.check(rc, "ds_rescale_inplace")
        ctx.dlogits_scale = g  # what the stored gradients carry now (a second backward reconciles against it)
        return dlogits.view(ctx.logits_shape), None, None, None, None, None


def weighted_ce(logits: torch.Tensor, target: torch.Tensor, class_weights: torch.Tensor, *, grad_scaler=None,
                class_counts: Optional[torch.Tensor] = None, return_stats: bool = False):
    """``nn.CrossEntropyLoss(class_weights)(logits, target)`` (main_deep_staple.py:716) with its
    backward (:719) fused into the same pass.  ``class_counts`` (``[C]`` label histogram of the batch)
    may be supplied to skip the label pre-pass (e.g. the all-reduced global histogram under data
    parallelism).  ``grad_scaler``: as for ``dp_weighted_loss`` -- pass it whenever AMP is on."""
    scale = _scale_tensor(grad_scaler, logits.device)
    if torch.is_grad_enabled() and logits.requires_grad:
        holder = []
        loss = _WeightedCE.apply(logits, target, class_weights, scale, class_counts, holder)
        arena = holder[0]
    else:
        z, t, B, C, V = _flat_inputs(logits, target)
        cw = _f32_vector(class_weights, z.device, C, "class weights for that many classes")
        arena = _launch_weighted_ce(z, t, B, C, V, cw, None, scale, None)
        loss = arena[_arena_layout(B, C)[1] + 1]
    if not return_stats:
        return loss
    B, C = logits.shape[0], logits.shape[1]
    return loss, arena[:_arena_layout(B, C)[0]].view(_F64).view(B, C, _b.DS_STATS)


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
