"""Oracle for the binary STAPLE baseline of the consensus notebook.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

The reference calls a third-party implementation that is absent from the reference tree and from this image:
``SimpleITK.STAPLEImageFilter`` (SimpleITK 2.1.1, ``poetry.lock:1592-1593``; ITK 5.2's
``itk::STAPLEImageFilter``) at ``deep_staple/consensus/consensus.ipynb:112-124``::

    staple_filter.SetForegroundValue(1.0); staple_filter.SetMaximumIterations(200)
    staple_out = staple_filter.Execute([label volumes as int16])
    consensus = staple_out > .5;  sensitivities = GetSensitivity();  specificities = GetSpecificity()

This file restates the PUBLISHED algorithm (Warfield, Zou, Wells: "Simultaneous truth and performance level
estimation (STAPLE)", IEEE TMI 23(7), 2004) in the form ITK's filter implements it:

* ``D_ij`` = 1 where rater j labels voxel i with the foreground value, else 0;
* initial estimate ``W_i`` = mean_j ``D_ij``; stationary prior ``g`` = confidence_weight * mean_i ``W_i``
  (confidence weight 1.0, the filter's default), kept fixed during the iterations;
* every iteration: M-step ``p_j = sum_i W_i D_ij / sum_i W_i`` (sensitivity),
  ``q_j = sum_i (1 - W_i)(1 - D_ij) / sum_i (1 - W_i)`` (specificity), then E-step
  ``W_i = g a_i / (g a_i + (1 - g) b_i)`` with ``a_i = prod_j (D_ij ? p_j : 1 - p_j)``,
  ``b_i = prod_j (D_ij ? 1 - q_j : q_j)``;
* stop after the E-step of the first iteration in which no ``p_j`` / ``q_j`` moved by more than 1e-7
  (squared change <= 1e-14), or after ``max_iterations``;
* outputs: the probability map ``W`` (thresholded at 0.5 by the notebook) and the final ``p``, ``q``.

There is no golden vector for it anywhere in the reference (the notebook's stored outputs are Dice printouts of
unpublished data), ITK cannot be installed here (no network), so this oracle is pinned only to the paper's equations:
"parity unpinned", see DESIGN.md.  Everything is float64 numpy.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

CONVERGENCE_SQ = 1.0e-14


def staple(labels: np.ndarray, foreground: float = 1.0, max_iterations: int = 200, confidence_weight: float = 1.0):
    """``labels [N, V]`` (any integer dtype) -> (W ``[V]`` float64, p ``[N]``, q ``[N]``, iterations run)."""
    d = (labels.reshape(labels.shape[0], -1) == foreground).astype(np.float64)
    n, v = d.shape
    w = d.mean(0)
    g = confidence_weight * w.mean()
    p = np.zeros(n)
    q = np.zeros(n)
    last_p = np.full(n, -10.0)
    last_q = np.full(n, -10.0)
    it = 0
    for it in range(1, max_iterations + 1):
        sw = w.sum()
        p = (d @ w) / sw
        q = ((1.0 - d) @ (1.0 - w)) / (v - sw)
        with np.errstate(divide="ignore", invalid="ignore"):
            a = np.prod(np.where(d > 0, p[:, None], 1.0 - p[:, None]), axis=0)
            b = np.prod(np.where(d > 0, 1.0 - q[:, None], q[:, None]), axis=0)
            w = g * a / (g * a + (1.0 - g) * b)
        done = bool(np.all((p - last_p) ** 2 <= CONVERGENCE_SQ) and np.all((q - last_q) ** 2 <= CONVERGENCE_SQ))
        last_p, last_q = p.copy(), q.copy()
        if done:
            break
    return w, p, q, it


def calc_staple_consensus(lbl_list: Sequence[torch.Tensor], max_iterations: int = 200) -> Tuple[torch.Tensor, List[float], List[float]]:
    """The notebook's ``calc_staple_consensus(lbl_list)`` (consensus.ipynb:112-124) on the restated filter:
    ``(consensus int64 [D,H,W], sensitivities, specificities)``."""
    dense = [(l.to_dense() if l.is_sparse else l) for l in lbl_list]
    shape = tuple(dense[0].shape)
    arr = np.stack([x.numpy().astype(np.int16).reshape(-1) for x in dense])  # the notebook casts to int16 (:118)
    w, p, q, _ = staple(arr, 1.0, max_iterations)
    consensus = torch.from_numpy((w > 0.5).reshape(shape)).long()
    return consensus, [float(x) for x in p], [float(x) for x in q]
