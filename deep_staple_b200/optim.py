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

__all__ = ["dp_optimizer_step"]


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
    scale = None
    if enabled:
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
