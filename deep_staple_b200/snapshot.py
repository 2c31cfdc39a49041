"""``train_label_snapshot.pth`` -- the on-disk contract between training and consensus.

Writer: main_deep_staple.py:963-1037 of the reference; reader: consensus.ipynb:330-350.  The
format is kept exactly (keys, dtypes, sparse-COO label stacks, ascending data-parameter order) so
snapshots written here load in the reference's notebook and vice versa.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import torch

__all__ = ["build_snapshot", "save_snapshot", "load_snapshot", "SNAPSHOT_KEYS"]

SNAPSHOT_KEYS = ("data_parameters", "disturb_flags", "d_ids", "dataset_idxs", "image_paths", "label_paths")
LABEL_KEYS = ("labels", "modified_labels", "train_predictions")


def _sparse(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_sparse else t.to_sparse()


def build_snapshot(
    data_parameters: torch.Tensor,
    d_ids: Sequence[str],
    dataset_idxs: Sequence,
    *,
    disturb_flags: Optional[Sequence[bool]] = None,
    image_paths: Optional[Sequence[str]] = None,
    label_paths: Optional[Sequence[str]] = None,
    labels: Optional[Sequence[torch.Tensor]] = None,
    modified_labels: Optional[Sequence[torch.Tensor]] = None,
    train_predictions: Optional[Sequence[torch.Tensor]] = None,
) -> Dict:
    """Assemble the dict of main_deep_staple.py:1010-1035 from per-instance entries.

    ``data_parameters`` ``[K]`` / ``[K,1]`` (``embedding(all_idxs)[train_idxs]``, :974-977); entries
    are sorted ascending by data parameter (:997); label tensors are stored as stacked sparse COO
    int64 (``labels``/``modified_labels`` ``[K,*S]``, ``train_predictions`` ``[K,1,*S]``)."""
    dp = data_parameters.detach().reshape(-1, 1).cpu()
    K = dp.shape[0]
    order = sorted(range(K), key=lambda i: float(dp[i, 0]))  # stable, like sorted(save_data, key=tpl[0])
    pick = lambda seq, default: tuple((seq[i] if seq is not None else default) for i in order)
    idxs = [torch.as_tensor(dataset_idxs[i]).reshape(()).long() for i in order]
    snap = {
        # the reference stacks rows of embedding(all_idxs) (:974-977, :1010): a non-leaf that still requires grad, and
        # torch.save / torch.load keep that flag (probed in SURVEY.md section 8 a10)
        "data_parameters": torch.stack([dp[i] for i in order]).requires_grad_(True),
        "disturb_flags": tuple(bool(x) for x in pick(disturb_flags, False)),
        "d_ids": pick(d_ids, ""),
        "dataset_idxs": torch.stack(idxs).cpu(),
        "image_paths": pick(image_paths, ""),
        "label_paths": pick(label_paths, ""),
    }
    if modified_labels is not None:
        snap["modified_labels"] = torch.stack([_sparse(modified_labels[i].long().cpu()) for i in order])
        src = labels if labels is not None else modified_labels
        snap["labels"] = torch.stack([_sparse(src[i].long().cpu()) for i in order])
        if train_predictions is not None:
            preds = [train_predictions[i].long().cpu() for i in order]
        else:
            preds = [torch.zeros_like(modified_labels[i].to_dense() if modified_labels[i].is_sparse else modified_labels[i]).long().cpu() for i in order]
        preds = [p.to_dense() if p.is_sparse else p for p in preds]
        preds = [p if p.dim() == snap["labels"].dim() else p.unsqueeze(0) for p in preds]
        snap["train_predictions"] = torch.stack([_sparse(p) for p in preds])
    return snap


def save_snapshot(snapshot: Dict, path) -> str:
    path = os.fspath(path)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(snapshot, path)  # :1037
    return path


def load_snapshot(path, map_location="cpu") -> Dict:
    """``torch.load`` of consensus.ipynb:344-350 plus a format check."""
    snap = torch.load(os.fspath(path), map_location=map_location)
    missing = [k for k in SNAPSHOT_KEYS if k not in snap]
    if missing:
        raise ValueError(f"not a train_label_snapshot: missing keys {missing}")
    return snap
