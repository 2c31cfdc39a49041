"""Hard Dice and the per-fold label pre-pass from the fused counting kernels.

Reference (multimodallearning/deep_staple):
* ``_diceNd``                      deep_staple/metrics.py:68-111  (called at main_deep_staple.py:782-786,
                                   consensus.ipynb:398-407)
* per-fold constants               main_deep_staple.py:646-653 (class weights, fixed weighting)

The reference one-hot encodes both label maps (8*C bytes per voxel) and loops over classes in
Python; here one streaming pass over the label maps (``ds_label_stats``) or the statistics the loss
pass already produced (``DPLossOutput.inst_stats``) give the same counts.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _binding as _b
from .loss import label_statistics

__all__ = ["dice_from_stats", "dice_from_label_maps", "prepass_constants"]


def dice_from_stats(inst_stats: torch.Tensor, nan_for_unlabeled_target: bool = True) -> torch.Tensor:
    """``2 TP / (pred + labelled)`` per instance and class from ``[B, C, 4]`` statistics -> CPU fp32
    ``[B, C]`` like the reference (metrics.py:89, 103-108: NaN for an empty denominator unless
    ``nan_for_unlabeled_target`` is False, then ``+1e-10``)."""
    tp = inst_stats[..., _b.DS_STAT_TP]
    den = inst_stats[..., _b.DS_STAT_PRED] + inst_stats[..., _b.DS_STAT_LABEL]
    if not nan_for_unlabeled_target:
        den = den + 1e-10
    return (2.0 * tp / den).to(torch.float32).cpu()


def dice_from_label_maps(pred: torch.Tensor, target: torch.Tensor, num_classes: int,
                         nan_for_unlabeled_target: bool = True) -> torch.Tensor:
    """Hard Dice of integer label maps ``[B, *spatial]`` (the arguments the call sites one-hot encode,
    main_deep_staple.py:782-786).  CPU inputs are moved to the current CUDA device."""
    if not torch.cuda.is_available():
        raise RuntimeError("deep_staple_b200 needs a B200 GPU; there is no CPU fallback")
    dev = pred.device if pred.is_cuda else (target.device if target.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    stats = label_statistics(pred.to(dev), target.to(dev), num_classes)
    return dice_from_stats(stats, nan_for_unlabeled_target)


def prepass_constants(mod_labels: torch.Tensor, num_classes: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Per-fold constants of main_deep_staple.py:646-653 from ALL modified labels ``[M, *spatial]`` in
    one launch: ``class_weights = count^-0.35 / mean`` (:650-651), ``fixed_weighting =
    log(gt_num + e) + e`` (:653) with ``gt_num`` = foreground voxels per instance (:648).
    Returns ``(class_weights [C], fixed_weighting [M], counts [C], gt_num [M])`` on the labels' device."""
    stats = label_statistics(mod_labels, None, num_classes)
    V = mod_labels[0].numel()
    per_class = stats[:, :, _b.DS_STAT_PRED]
    counts = per_class.sum(0).to(torch.float32)
    gt_num = (V - per_class[:, 0]).to(torch.float32)
    class_weights = 1 / counts.pow(0.35)
    class_weights = class_weights / class_weights.mean()
    fixed_weighting = (gt_num + math.e).log() + math.e
    return class_weights, fixed_weighting, counts, gt_num
