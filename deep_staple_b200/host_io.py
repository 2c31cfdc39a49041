"""Upload path for label maps that live on the HOST as int64 (the reference's container dtype).

Replaces ``b_seg_modified.to(device=config.device)`` (main_deep_staple.py:687) and the dense int64
candidate uploads of the consensus notebook (consensus.ipynb:382-392): the labels are packed to uint8 on
the host by ``ds_host_narrow_labels`` (C ABI, worker threads) into a pinned staging buffer and THAT is
copied -- 1 byte per voxel over PCIe instead of 8.  Every kernel of this package reads uint8 labels
natively and returns bit-identical results for them (tests/test_gpu_loss.py, tests/test_gpu_consensus.py).
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _binding as _b

_MAX_THREADS = 16


def default_threads() -> int:
    """Worker threads of the packer: the CPUs this process may run on (its NUMA binding), at most 16, and no more
    than this rank's share of the machine when several ranks run on it (torchrun's LOCAL_WORLD_SIZE)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        n = os.cpu_count() or 1
    try:
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    except ValueError:
        local_world = 1
    share = max(1, (os.cpu_count() or n) // local_world)
    return max(1, min(n, share, _MAX_THREADS))


def narrow_labels_host(labels: torch.Tensor, out: Optional[torch.Tensor] = None, threads: Optional[int] = None) -> torch.Tensor:
    """int64 host label map -> uint8 host label map of the same shape (``out``: typically pinned).

    Raises ``ValueError`` if a label lies outside [0, 255] (nothing is silently truncated)."""
    if labels.is_cuda or labels.dtype != torch.int64:
        raise ValueError("narrow_labels_host takes an int64 tensor in host memory")
    src = labels.contiguous()
    if out is None:
        out = torch.empty(src.shape, dtype=torch.uint8)
    elif out.is_cuda or out.dtype != torch.uint8 or out.numel() != src.numel() or not out.is_contiguous():
        raise ValueError("out must be a contiguous uint8 host tensor with as many elements as labels")
    if src.numel() == 0:
        return out
    rc = _b.load().ds_host_narrow_labels(src.data_ptr(), out.data_ptr(), src.numel(), threads or default_threads())
    _b.check(rc, "ds_host_narrow_labels")
    return out


class LabelUploader:
    """Double-buffered host -> device label upload: pack (host threads) into a pinned staging slot, then one
    asynchronous copy of the packed bytes on the caller's current stream.

        up = ds.LabelUploader(labels.shape, device)          # once
        d_labels = up(labels_int64_host)                     # per step, instead of labels.to(device)

    A staging slot is reused only after the copy that last read it has completed (event per slot), so the
    packer of step k + 1 runs while the copy of step k is in flight.  The returned uint8 device tensor is
    owned by the uploader and valid until the call after next (``slots`` calls later).  uint8 host labels (packed
    earlier, e.g. in a data-loader worker) are copied as they are."""

    def __init__(self, shape, device, slots: int = 2, threads: Optional[int] = None):
        self.device = torch.device(device)
        self.shape = tuple(shape)
        self.threads = threads or default_threads()
        self._staging = [torch.empty(self.shape, dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self._device = [torch.empty(self.shape, dtype=torch.uint8, device=self.device) for _ in range(slots)]
        self._copied = [None] * slots
        self._i = 0

    @property
    def h2d_bytes(self) -> int:
        return self._staging[0].numel()

    def __call__(self, labels: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        k = self._i % len(self._staging)
        self._i += 1
        if self._copied[k] is not None:
            self._copied[k].synchronize()  # the previous copy out of this slot
        dst = self._device[k] if out is None else out
        if labels.dtype is torch.uint8 and not labels.is_cuda:
            # packed already (e.g. by ``narrow_labels_host`` in a data-loader worker): nothing to do on this thread
            # but the copy -- straight out of the caller's tensor when it is pinned, through the staging slot otherwise
            if tuple(labels.shape) != self.shape:
                raise ValueError(f"labels have shape {tuple(labels.shape)}, the uploader was built for {self.shape}")
            src = labels if labels.is_contiguous() else labels.contiguous()
            if not src.is_pinned():
                self._staging[k].copy_(src)
                src = self._staging[k]
            dst.copy_(src, non_blocking=True)
        else:
            narrow_labels_host(labels, self._staging[k], self.threads)
            dst.copy_(self._staging[k], non_blocking=True)
        ev = self._copied[k] or torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._copied[k] = ev
        return dst
