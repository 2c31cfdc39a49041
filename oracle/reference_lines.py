"""Run the reference's OWN source lines for the hot path.  TEST INFRASTRUCTURE ONLY.

Used by ``tests/golden/make_golden.py`` (golden-vector generation, in the build container where
``/root/reference`` is mounted read-only) and by the drift-guard tests.  Nothing is copied into
this repo: the text is read from the reference tree at call time, dedented and ``exec``-ed with the
surrounding variables of ``train_DL`` supplied by the caller.  On the GPU box the reference tree
does not exist; ``available()`` is False there and callers must use the committed fixtures.
"""
from __future__ import annotations

import importlib.util
import json
import os
import textwrap
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("DEEP_STAPLE_REFERENCE", "/root/reference")
MAIN = "main_deep_staple.py"
NOTEBOOK = "deep_staple/consensus/consensus.ipynb"

# (first line, last line, anchor text that must appear in the first line) -- drift guard
LOSS_CE_LINES = (716, 716, "nn.CrossEntropyLoss(class_weights)(logits, b_seg_modified)")
LOSS_CE_BACKWARD_LINES = (719, 719, "scaler.scale(ce_loss).backward()")
LOSS_DP_LINES = (738, 759, "nn.CrossEntropyLoss(reduction='none')(dp_logits, b_seg_modified)")
LOSS_BACKWARD_LINES = (762, 762, "scaler_dp.scale(dp_loss).backward()")
PREPASS_BINCOUNT_LINE = (646, 646, "torch.bincount(mod_label.reshape(-1).long()")
PREPASS_GTNUM_LINE = (648, 648, "(mod_label > 0).sum(dim=n_dims)")
PREPASS_CONST_LINES = (650, 653, "class_weights = 1 / (bn_count).float().pow(.35)")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, MAIN)) and os.path.isfile(
        os.path.join(REFERENCE_ROOT, NOTEBOOK)
    )


def _main_lines(span) -> str:
    first, last, anchor = span
    with open(os.path.join(REFERENCE_ROOT, MAIN)) as fh:
        lines = fh.read().split("\n")
    block = lines[first - 1 : last]
    if anchor not in block[0]:
        raise RuntimeError(f"reference drift: {MAIN}:{first} no longer contains {anchor!r}")
    return textwrap.dedent("\n".join(block)) + "\n"


def consensus_source() -> str:
    """Source text of ``calc_dp_consensus`` (consensus.ipynb, the cell that defines it, cut before
    ``calc_staple_consensus`` which needs SimpleITK)."""
    with open(os.path.join(REFERENCE_ROOT, NOTEBOOK)) as fh:
        nb = json.load(fh)
    for cell in nb["cells"]:
        src = "".join(cell.get("source", []))
        if cell.get("cell_type") == "code" and "def calc_dp_consensus" in src:
            start = src.index("def calc_dp_consensus")
            end = src.index("def calc_staple_consensus")
            return src[start:end]
    raise RuntimeError("reference drift: calc_dp_consensus not found in the consensus notebook")


def reference_calc_dp_consensus():
    ns = {"torch": torch}
    exec(compile(consensus_source(), "consensus.ipynb:calc_dp_consensus", "exec"), ns)
    return ns["calc_dp_consensus"]


def reference_metrics_module():
    path = os.path.join(REFERENCE_ROOT, "deep_staple/metrics.py")
    spec = importlib.util.spec_from_file_location("_reference_deep_staple_metrics", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Scaler:
    """Stand-in for ``amp.GradScaler``: ``scale(x) = x * g`` (the real one disables itself on CPU)."""

    def __init__(self, g: float):
        self.g = g

    def scale(self, x):
        return x * self.g


def reference_loss(
    logits: torch.Tensor,
    target: torch.Tensor,
    embedding_weight: torch.Tensor,
    idx: torch.Tensor,
    fixed_weighting: Optional[torch.Tensor],
    class_weights: torch.Tensor,
    use_risk_regularization: bool,
    use_fixed_weighting: bool,
    grad_scale: float = 1.0,
    need_dlogits: bool = True,
):
    """Execute main_deep_staple.py:716, :738-759 and :762 verbatim on the given tensors.

    Half-precision logits are fed through an explicit ``.float()`` (with autograd through the cast):
    that is what CUDA autocast does to ``cross_entropy`` inputs (main_deep_staple.py:701), so the
    gradient comes back rounded to the logits dtype exactly as in the reference's AMP run."""
    m = embedding_weight.shape[0]
    embedding = nn.Embedding(m, 1, sparse=True, _weight=embedding_weight.detach().clone().reshape(m, 1))
    z_leaf = logits.detach().clone().requires_grad_(need_dlogits)
    z = z_leaf.float() if z_leaf.dtype in (torch.float16, torch.bfloat16) else z_leaf
    use_2d = logits.dim() == 4
    ns = {
        "nn": nn,
        "torch": torch,
        "np": np,
        "config": SimpleNamespace(
            use_fixed_weighting=use_fixed_weighting,
            use_risk_regularization=use_risk_regularization,
            use_2d_normal_to=("D" if use_2d else None),
        ),
        "logits": z,
        "dp_logits": z,
        "b_seg_modified": target,
        "b_idxs_dataset": idx,
        "class_weights": class_weights,
        "embedding": embedding,
        "fixed_weighting": fixed_weighting,
        "n_dims": (-2, -1) if use_2d else (-3, -2, -1),
        "scaler_dp": _Scaler(grad_scale),
    }
    exec(compile(_main_lines(LOSS_CE_LINES), f"{MAIN}:716", "exec"), ns)
    exec(compile(_main_lines(LOSS_DP_LINES), f"{MAIN}:738-759", "exec"), ns)
    exec(compile(_main_lines(LOSS_BACKWARD_LINES), f"{MAIN}:762", "exec"), ns)
    grad = embedding.weight.grad.coalesce()

    # class-weighted CE with its own backward (:716 + :719) on a fresh leaf
    z_ce_leaf = logits.detach().clone().requires_grad_(True)
    z_ce = z_ce_leaf.float() if z_ce_leaf.dtype in (torch.float16, torch.bfloat16) else z_ce_leaf
    ns_ce = dict(ns, logits=z_ce, scaler=_Scaler(grad_scale))
    exec(compile(_main_lines(LOSS_CE_LINES), f"{MAIN}:716", "exec"), ns_ce)
    exec(compile(_main_lines(LOSS_CE_BACKWARD_LINES), f"{MAIN}:719", "exec"), ns_ce)
    return {
        "ce_loss": ns["ce_loss"].detach(),
        "dp_loss": ns["dp_loss"].detach(),
        "d_bare_weight": grad.to_dense().reshape(-1)[idx.reshape(-1)],
        "d_embedding_is_sparse": bool(embedding.weight.grad.is_sparse),
        "d_logits": z_leaf.grad if need_dlogits else None,
        "d_logits_ce": z_ce_leaf.grad,
    }


def reference_prepass(mod_labels, num_classes: int):
    """Execute main_deep_staple.py:646, :648 per sample and :650-653 once."""
    ns = {
        "torch": torch,
        "np": np,
        "training_dataset": SimpleNamespace(label_tags=tuple(range(num_classes))),
        "bn_count": torch.zeros([num_classes]),
        "gt_num": torch.zeros([len(mod_labels)]),
    }
    line_bin = _main_lines(PREPASS_BINCOUNT_LINE)
    line_gt = _main_lines(PREPASS_GTNUM_LINE)
    for i, lbl in enumerate(mod_labels):
        ns["mod_label"] = lbl
        ns["d_idxs"] = i
        ns["n_dims"] = tuple(range(-lbl.dim(), 0))
        exec(compile(line_bin, f"{MAIN}:646", "exec"), ns)
        exec(compile(line_gt, f"{MAIN}:648", "exec"), ns)
    exec(compile(_main_lines(PREPASS_CONST_LINES), f"{MAIN}:650-653", "exec"), ns)
    return ns["class_weights"], ns["fixed_weighting"], ns["bn_count"], ns["gt_num"]
