"""Oracle for the hard Dice metric.  TEST INFRASTRUCTURE ONLY.

Follows deep_staple/metrics.py:68-111 (``_diceNd``; ``dice2d`` :7-31 and ``dice3d`` :39-60 only
assert the rank and forward to it)."""
from __future__ import annotations

import torch


def dice_nd(predicted, target, one_hot_torch_style: bool, nan_for_unlabeled_target: bool = True):
    """Per-instance per-class ``2 TP / (pred + labelled)`` on one-hot inputs -> CPU fp32 ``[B,C]``.
    ``nan_for_unlabeled_target=False`` adds 1e-10 to the denominator (metrics.py:103-106)."""
    if one_hot_torch_style:  # metrics.py:84-86
        predicted = torch.movedim(predicted, -1, 1)
        target = torch.movedim(target, -1, 1)
    b, c = target.shape[0], target.shape[1]
    out = torch.zeros((b, c))
    eps = 0.0 if nan_for_unlabeled_target else 1e-10
    for k in range(c):  # metrics.py:93-108
        pred_k = predicted[:, k] == 1
        lab_k = target[:, k] == 1
        tp = torch.logical_and(pred_k, lab_k).reshape(b, -1).sum(-1)
        n_pred = pred_k.reshape(b, -1).sum(-1)
        n_lab = lab_k.reshape(b, -1).sum(-1)
        out[:, k] = 2.0 * tp / (n_pred + n_lab + eps)
    return out


def dice_from_label_maps(pred_lbl, target_lbl, num_classes: int, nan_for_unlabeled_target: bool = True):
    """The call pattern of main_deep_staple.py:782-786: one-hot both integer maps torch-style."""
    oh = torch.nn.functional.one_hot
    return dice_nd(oh(pred_lbl.long(), num_classes), oh(target_lbl.long(), num_classes), True, nan_for_unlabeled_target)
