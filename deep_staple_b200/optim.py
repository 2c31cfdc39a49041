"""Fused data-parameter optimiser step -- `scaler_dp.step(optimizer_dp)` of main_deep_staple.py:764-773
(GradScaler unscale + inf check + SparseAdam on the touched rows) as ONE kernel launch, keeping
torch's own objects as the owners of all state: `optimizer_dp.state_dict()`, `scaler_dp.state_dict()`
and `embedding.weight` stay interchangeable with the reference's checkpoints (torch_utils.py:354-360).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _binding as _b

__all__ = ["dp_optimizer_step", "LossSection"]

import os


def _check_unique(idx: torch.Tensor) -> None:
    """``DS_DEBUG=1``: the optimiser kernels assume unique rows per batch (the reference's ``SubsetRandomSampler``,
    main_deep_staple.py:597, guarantees it; torch's SparseAdam would coalesce duplicates, the kernels would race)."""
    if os.environ.get("DS_DEBUG") and idx.unique().numel() != idx.numel():
        raise ValueError("b_idxs_dataset contains duplicate rows: the fused data-parameter step needs unique indices")


def _require_scaler_internals(scaler) -> None:
    """The fused steps read / write private GradScaler state; fail loudly if a torch upgrade moved it."""
    for attr in ("_scale", "_growth_tracker", "_per_optimizer_states", "_lazy_init_scale_growth_tracker",
                 "get_growth_factor", "get_backoff_factor", "get_growth_interval"):
        if not hasattr(scaler, attr):
            raise RuntimeError(f"torch.amp.GradScaler has no attribute {attr!r} in torch {torch.__version__}: "
                               "deep_staple_b200's fused optimiser step does not support this torch version")


def dp_optimizer_step(
    embedding: torch.nn.Embedding,
    optimizer_dp: torch.optim.SparseAdam,
    b_idxs_dataset: torch.Tensor,
    d_bare_weight: torch.Tensor,
    scaler_dp=None,
    *,
    sync: bool = True,
) -> Optional[bool]:
    """Replacement for ``scaler_dp.step(optimizer_dp)`` (:767); call ``scaler_dp.update()`` afterwards
    exactly like the reference (:768) -- the scaler sees the inf flag of this step.

    embedding       ``nn.Embedding(M, 1, sparse=True)`` (:418); its weight is updated in place
    optimizer_dp    the ``SparseAdam`` that owns the moments (:442-444); its state is used and updated
    b_idxs_dataset  ``[B]`` rows of this batch (unique, as the reference's sampler guarantees)
    d_bare_weight   ``[B]`` gradient of the SCALED loss w.r.t. the drawn data parameters -- the
                    ``d_bare_weight`` field of ``dp_weighted_loss(..., return_all=True)`` -- so no
                    autograd backward is needed for the data parameters at all
    scaler_dp       the ``GradScaler`` whose scale multiplied the loss (``None`` / disabled: scale 1)
    sync            read the inf flag back (one host sync, which ``GradScaler.step`` does too) so that the
                    optimiser's step counter only advances when the step was taken.  ``sync=False``
                    advances it unconditionally (the kernel still skips the update on inf / nan).
    returns         True if the step was applied, False if skipped (None with ``sync=False``)."""
    w = embedding.weight
    _b.require_cuda(w, "embedding.weight")
    if w.dim() != 2 or w.shape[1] != 1 or w.dtype != torch.float32 or not w.is_contiguous():
        raise ValueError("the data-parameter embedding must be a contiguous float32 [M, 1] table")
    group = next(g for g in optimizer_dp.param_groups if any(p is w for p in g["params"]))
    if group.get("maximize", False):
        raise ValueError("maximize=True is not supported")
    state = optimizer_dp.state[w]
    if len(state) == 0:  # same lazy initialisation as SparseAdam.step
        state["step"] = 0
        state["exp_avg"] = torch.zeros_like(w, memory_format=torch.preserve_format)
        state["exp_avg_sq"] = torch.zeros_like(w, memory_format=torch.preserve_format)
    beta1, beta2 = group["betas"]
    lr = float(group["lr"])
    step = int(state["step"]) + 1
    step_size = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    dev = w.device
    idx = b_idxs_dataset.detach().to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
    grad = d_bare_weight.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
    if idx.numel() != grad.numel():
        raise ValueError(f"{idx.numel()} indices but {grad.numel()} gradients")
    enabled = scaler_dp is not None and scaler_dp.is_enabled()
    _check_unique(idx)
    scale = None
    if enabled:
        _require_scaler_internals(scaler_dp)
        scale = scaler_dp._scale
        if scale is None:
            raise RuntimeError("the GradScaler has not scaled a loss yet: pass the same scaler to dp_weighted_loss(grad_scaler=...)")
    found_inf = torch.zeros(1, dtype=torch.float32, device=dev)
    lib = _b.load()
    with torch.no_grad(), torch.cuda.device(dev):
        rc = lib.ds_sparse_adam_step(
            _b.ptr(w), _b.ptr(state["exp_avg"]), _b.ptr(state["exp_avg_sq"]), _b.ptr(idx), _b.ptr(grad), idx.numel(),
            w.shape[0], _b.ptr(scale), _b.ptr(found_inf), float(beta1), float(beta2), float(group["eps"]), step_size,
            _b.stream_ptr(dev),
        )
    _b.check(rc, "ds_sparse_adam_step")
    if enabled:
        # hand the inf flag to GradScaler.update() exactly where GradScaler.step() would have left it
        from torch.amp.grad_scaler import OptState

        st = scaler_dp._per_optimizer_states[id(optimizer_dp)]
        st["found_inf_per_device"] = {dev: found_inf}
        st["stage"] = OptState.STEPPED
    applied = None
    if sync:
        applied = not bool(found_inf.item())
        if applied:
            state["step"] = step
    else:
        state["step"] = step
    return applied


class LossSection:
    """The out-of-line data-parameter step of the reference's training loop (``use_ool_dp_loss=True``, its
    default; main_deep_staple.py:738-773) as TWO kernel launches and nothing else:

        dp_loss value + its gradient w.r.t. the drawn data parameters     (:738-759, :762)   ds_dp_loss_fwd_bwd_indexed
        scaler_dp.step(optimizer_dp); scaler_dp.update()                  (:764-768)         ds_sparse_adam_scaler_step

    No autograd graph, no ``embedding(idx)`` / ``fixed_weighting[idx]`` gathers (the kernel draws both by dataset
    index), no sparse gradient tensor, no host synchronisation -- the host cost of a call is two ctypes calls, and
    because nothing depends on host state the call can be captured in a CUDA graph when its inputs are static.
    torch's objects stay the owners of all state: ``embedding.weight``, ``optimizer_dp.state`` (moments; the step
    counter is mirrored on the device and written back by ``sync()`` and automatically before
    ``optimizer_dp.state_dict()``), ``scaler_dp`` (its ``_scale`` / ``_growth_tracker`` tensors are updated in place),
    so checkpoints (torch_utils.py:354-360) are unchanged.

        section = ds.LossSection(embedding, optimizer_dp, scaler_dp, fixed_weighting=fixed_weighting)
        dp_loss = section(dp_logits, b_seg_modified, b_idxs_dataset)        # replaces :738-768 in the OOL branch
    """

    def __init__(self, embedding: torch.nn.Embedding, optimizer_dp: torch.optim.SparseAdam, scaler_dp=None, *,
                 fixed_weighting: Optional[torch.Tensor] = None, use_risk_regularization: bool = True,
                 class_weights: Optional[torch.Tensor] = None):
        w = embedding.weight
        _b.require_cuda(w, "embedding.weight")
        if w.dim() != 2 or w.shape[1] != 1 or w.dtype != torch.float32 or not w.is_contiguous():
            raise ValueError("the data-parameter embedding must be a contiguous float32 [M, 1] table")
        self.embedding, self.optimizer, self.scaler = embedding, optimizer_dp, scaler_dp
        self.group = next(g for g in optimizer_dp.param_groups if any(p is w for p in g["params"]))
        if self.group.get("maximize", False):
            raise ValueError("maximize=True is not supported")
        dev = w.device
        self.device = dev
        state = optimizer_dp.state[w]
        if len(state) == 0:  # same lazy initialisation as SparseAdam.step
            state["step"] = 0
            state["exp_avg"] = torch.zeros_like(w, memory_format=torch.preserve_format)
            state["exp_avg_sq"] = torch.zeros_like(w, memory_format=torch.preserve_format)
        self.state = state
        self.step_count = torch.tensor([int(state["step"])], dtype=torch.int64, device=dev)
        self.found_inf = torch.zeros(1, dtype=torch.float32, device=dev)
        self.fixed = None
        if fixed_weighting is not None:
            self.fixed = fixed_weighting.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
            if self.fixed.numel() != w.shape[0]:
                raise ValueError("fixed_weighting must have one entry per embedding row")
        self.cw = None if class_weights is None else class_weights.detach().to(device=dev, dtype=torch.float32).contiguous()
        self.use_risk = 1 if use_risk_regularization else 0
        self.scale = self.tracker = None
        self.growth = (2.0, 0.5, 2000)
        if scaler_dp is not None and scaler_dp.is_enabled():
            _require_scaler_internals(scaler_dp)  # private state: fail loudly instead of silently mis-scaling
            if scaler_dp._scale is None:
                scaler_dp._lazy_init_scale_growth_tracker(dev)
            self.scale, self.tracker = scaler_dp._scale, scaler_dp._growth_tracker
            if self.scale.dtype != torch.float32 or self.tracker.dtype != torch.int32:
                raise RuntimeError("unexpected GradScaler tensor dtypes: unsupported torch version")
            self.growth = (float(scaler_dp.get_growth_factor()), float(scaler_dp.get_backoff_factor()),
                           int(scaler_dp.get_growth_interval()))
        if hasattr(optimizer_dp, "register_state_dict_pre_hook"):
            optimizer_dp.register_state_dict_pre_hook(lambda opt: self.sync())
        self._lib = _b.load()

    def sync(self) -> int:
        """Write the device-side step counter back into ``optimizer_dp.state`` (one device->host read)."""
        self.state["step"] = int(self.step_count.item())
        return self.state["step"]

    def __call__(self, dp_logits: torch.Tensor, b_seg_modified: torch.Tensor, b_idxs_dataset: torch.Tensor,
                 *, dice_target: Optional[torch.Tensor] = None, return_all: bool = False):
        from . import loss as _l

        z, t, B, C, V = _l._flat_inputs(dp_logits, b_seg_modified)
        dev = self.device
        idx = b_idxs_dataset
        if idx.dtype is not torch.int64 or idx.device != dev or idx.dim() != 1 or not idx.is_contiguous():
            idx = idx.detach().to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
        if idx.shape[0] != B:
            raise ValueError(f"{B} instances in the batch but {idx.shape[0]} dataset indices")
        _check_unique(idx)
        n_stats, o_loss, o_dbare, o_inst, total = _l._arena_layout(B, C)
        arena = torch.empty(total, dtype=torch.float32, device=dev)
        base = arena.data_ptr()
        ws = _l._workspace(dev, B, C)
        w, lib, stream = self.embedding.weight, self._lib, _b.stream_ptr(dev)
        beta1, beta2 = self.group["betas"]
        with torch.cuda.device(dev):
            rc = lib.ds_dp_loss_fwd_bwd_indexed(
                z.data_ptr(), _b.logit_dtype(z), t.data_ptr(), _b.label_dtype(t), _b.ptr(dice_target), B, C, V, w.data_ptr(),
                _b.ptr(self.fixed), idx.data_ptr(), _b.ptr(self.cw), self.use_risk, _b.ptr(self.scale), None, base,
                base + 4 * o_loss, base + 4 * o_dbare, base + 4 * o_inst, ws.data_ptr(), ws.numel() * 8, stream)
            if rc:
                _b.check(rc, "ds_dp_loss_fwd_bwd_indexed")
            rc = lib.ds_sparse_adam_scaler_step(
                w.data_ptr(), self.state["exp_avg"].data_ptr(), self.state["exp_avg_sq"].data_ptr(), idx.data_ptr(),
                base + 4 * o_dbare, B, w.shape[0], _b.ptr(self.scale), _b.ptr(self.tracker), self.step_count.data_ptr(),
                self.found_inf.data_ptr(), float(self.group["lr"]), float(beta1), float(beta2), float(self.group["eps"]),
                self.growth[0], self.growth[1], self.growth[2], stream)
            if rc:
                _b.check(rc, "ds_sparse_adam_scaler_step")
        loss = arena[o_loss]
        return _l.DPLossOutput(loss, arena, B, C) if return_all else loss
