"""Oracle for the data-parameter-weighted segmentation loss.  TEST INFRASTRUCTURE ONLY.

Restates the inline code of ``train_DL`` in the reference (it is not a function there and the
module starts training on import, so it cannot be imported):

* class-weighted CE for the network      main_deep_staple.py:716
* per-instance CE mean                   main_deep_staple.py:738-739
* data-parameter draw                    main_deep_staple.py:741
* sigmoid + batch-mean normalisation     main_deep_staple.py:743-744
* fixed weighting                        main_deep_staple.py:747-748
* risk regularisation                    main_deep_staple.py:750-757 (759 without)
* scaled backward                        main_deep_staple.py:762

All arithmetic is torch CPU ATen, fp32 by default (fp64 when the inputs are fp64 -> arbiter).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import math
import torch
from torch import nn


def spatial_dims(logits: torch.Tensor) -> Tuple[int, ...]:
    """main_deep_staple.py:513-516: ``(-2,-1)`` in 2D slice mode, ``(-3,-2,-1)`` in 3D."""
    return tuple(range(-(logits.dim() - 2), 0))


def class_weighted_ce(logits, target, class_weights) -> torch.Tensor:
    """main_deep_staple.py:716."""
    return nn.CrossEntropyLoss(class_weights)(logits, target)


def dp_loss_forward(
    dp_logits: torch.Tensor,
    target: torch.Tensor,
    bare_weight: torch.Tensor,
    fixed_weighting_b: Optional[torch.Tensor],
    use_risk_regularization: bool,
) -> torch.Tensor:
    """main_deep_staple.py:738-759 given the already drawn data parameters ``bare_weight [B]``
    (:741) and the already gathered ``fixed_weighting[b_idxs_dataset]`` (``None`` = flag off)."""
    n_dims = spatial_dims(dp_logits)
    per_inst = nn.CrossEntropyLoss(reduction="none")(dp_logits, target)  # :738
    per_inst = per_inst.mean(n_dims)  # :739
    weight = torch.sigmoid(bare_weight)  # :743
    weight = weight / weight.mean()  # :744
    if fixed_weighting_b is not None:
        weight = weight / fixed_weighting_b  # :748
    if use_risk_regularization:
        fg_pred = (dp_logits.argmax(1) > 0).sum(dim=n_dims).detach()  # :751
        numel = 1
        for d in n_dims:
            numel *= dp_logits.shape[d]
        risk = -weight * fg_pred / numel  # :753 / :755
        return (per_inst * weight).sum() + risk.sum()  # :757
    return (per_inst * weight).sum()  # :759


@dataclass
class LossResult:
    ce_loss: torch.Tensor
    dp_loss: torch.Tensor
    d_bare_weight: torch.Tensor  # [B] dense view of the sparse embedding gradient rows
    d_embedding: Optional[torch.Tensor]  # sparse COO [M,1] grad when an embedding was used
    d_logits: Optional[torch.Tensor]  # grad of dp_loss wrt dp_logits (non-OOL mode), logits dtype
    d_logits_ce: Optional[torch.Tensor]  # grad of ce_loss wrt logits
    per_inst_ce: torch.Tensor  # [B]  (:739)
    fg_pred: torch.Tensor  # [B] int64 (:751)


def run_loss(
    logits: torch.Tensor,
    target: torch.Tensor,
    embedding_weight: torch.Tensor,
    idx: torch.Tensor,
    fixed_weighting: Optional[torch.Tensor],
    class_weights: Optional[torch.Tensor],
    use_risk_regularization: bool = True,
    grad_scale: float = 1.0,
    need_dlogits: bool = True,
) -> LossResult:
    """One batch of the reference's loss section incl. backward (:716, :738-762) on CPU.

    ``embedding_weight`` is the ``[M,1]`` table of ``nn.Embedding(M, 1, sparse=True)`` (:418),
    ``idx`` the batch's ``b_idxs_dataset``; ``fixed_weighting`` the ``[M]`` vector (:653) or None.
    ``grad_scale`` stands in for ``scaler_dp.scale`` (:762; 1.0 on CPU where GradScaler disables
    itself).  ``need_dlogits=False`` is the out-of-line mode (:724-729: no grad to the logits)."""
    m = embedding_weight.shape[0]
    emb = nn.Embedding(m, 1, sparse=True, _weight=embedding_weight.detach().clone().reshape(m, 1))
    z = logits.detach().clone().requires_grad_(need_dlogits)
    tgt = target.long()

    z_ce = logits.detach().clone().requires_grad_(True)
    if class_weights is not None:
        ce = class_weighted_ce(z_ce.float() if z_ce.dtype in (torch.float16, torch.bfloat16) else z_ce, tgt, class_weights)
        (ce * grad_scale).backward()
        d_ce = z_ce.grad
    else:
        ce = torch.zeros(())
        d_ce = None

    # autocast runs cross_entropy in fp32 (main_deep_staple.py:701); CPU half CE is emulated by the
    # same explicit up-cast so that the gradient comes back in the logits dtype.
    z_in = z.float() if z.dtype in (torch.float16, torch.bfloat16) else z
    bare = emb(idx).squeeze()  # :741
    fixed_b = fixed_weighting[idx] if fixed_weighting is not None else None
    if fixed_b is not None and bare.dim() == 0:
        fixed_b = fixed_b.squeeze()
    loss = dp_loss_forward(z_in, tgt, bare, fixed_b, use_risk_regularization)
    (loss * grad_scale).backward()  # :762

    g = emb.weight.grad.coalesce()
    dense_rows = g.to_dense().reshape(-1)[idx.reshape(-1)]
    with torch.no_grad():
        n_dims = spatial_dims(z_in)
        per_inst = nn.CrossEntropyLoss(reduction="none")(z_in, tgt).mean(n_dims)
        fg_pred = (z_in.argmax(1) > 0).sum(dim=n_dims)
    return LossResult(
        ce_loss=ce.detach(),
        dp_loss=loss.detach(),
        d_bare_weight=dense_rows,
        d_embedding=g,
        d_logits=z.grad if need_dlogits else None,
        d_logits_ce=d_ce,
        per_inst_ce=per_inst,
        fg_pred=fg_pred,
    )


def closed_form(
    per_inst_ce: torch.Tensor,
    fg_pred: torch.Tensor,
    bare_weight: torch.Tensor,
    fixed_b: Optional[torch.Tensor],
    numel_spatial: int,
    use_risk_regularization: bool,
    grad_scale: float = 1.0,
):
    """SURVEY.md Appendix A in fp64: loss and d loss / d bare_weight from the per-instance
    statistics.  ``L = sum_b w_b (l_b - n_b/V)``, ``dL/dp_k = s_k (1-s_k) (c_k - L/B) / m``."""
    p = bare_weight.double().reshape(-1)
    b = p.numel()
    s = torch.sigmoid(p)
    mean_s = s.mean()
    f = fixed_b.double().reshape(-1) if fixed_b is not None else torch.ones_like(p)
    risk = fg_pred.double().reshape(-1) / numel_spatial if use_risk_regularization else torch.zeros_like(p)
    c = (per_inst_ce.double().reshape(-1) - risk) / f
    loss = (s * c).sum() / mean_s
    grad = grad_scale * s * (1 - s) * (c - loss / b) / mean_s
    weight = s / (mean_s * f)
    return loss, grad, weight


def prepass_constants(mod_labels, num_classes: int):
    """main_deep_staple.py:646-653: class weights ``1/count^0.35`` normalised by their mean and the
    fixed weighting ``log(gt_num + e) + e`` with ``gt_num`` = foreground voxels per instance."""
    counts = torch.zeros(num_classes)
    gt_num = torch.zeros(len(mod_labels))
    for i, lbl in enumerate(mod_labels):
        counts += torch.bincount(lbl.reshape(-1).long(), minlength=num_classes)  # :646
        gt_num[i] = (lbl > 0).sum().float()  # :648
    class_weights = 1 / counts.float().pow(0.35)  # :650
    class_weights /= class_weights.mean()  # :651
    fixed_weighting = (gt_num + math.e).log() + math.e  # :653
    return class_weights, fixed_weighting, counts, gt_num
