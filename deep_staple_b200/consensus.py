"""Consensus weighted vote on B200 -- the Python call surface of the reference, backed by the
sm_100a kernels behind ``ds_consensus_vote*`` (include/deep_staple_b200.h).

Reference (multimodallearning/deep_staple):
* ``calc_dp_consensus(lbl_list, weighting_list)``   deep_staple/consensus/consensus.ipynb:100-109
* the consensus driver cell                         deep_staple/consensus/consensus.ipynb:353-422

Everything here runs on the GPU; tensors that arrive on the CPU (the notebook works with
``DEVICE='cpu'``, consensus.ipynb:30) are copied to the current CUDA device, voted on there and the
result is returned on the device the labels came from.  There is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from . import _binding as _b

__all__ = [
    "calc_dp_consensus",
    "consensus_batched",
    "consensus_sparse",
    "consensus_multiclass",
    "calc_staple_consensus",
    "staple_batched",
    "consensus_from_snapshot",
    "group_snapshot",
]


def _cuda_device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("deep_staple_b200 needs a B200 (sm_100a) GPU: no CUDA device is visible and there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"device {device} requested: deep_staple_b200 has no CPU fallback")
    return device if device.index is not None else torch.device("cuda", torch.cuda.current_device())


def _weights_tensor(weighting_list, device) -> torch.Tensor:
    """``torch.tensor(weighting_list)`` of consensus.ipynb:103 as an fp32 ``[N]`` device tensor
    (accepts a tensor, or a list of floats / 0-dim / ``[1]`` tensors like the notebook passes)."""
    if isinstance(weighting_list, torch.Tensor):
        w = weighting_list.detach().reshape(-1)
    else:
        items = list(weighting_list)
        if len(items) and isinstance(items[0], torch.Tensor):
            w = torch.stack([x.detach().reshape(()) for x in items])
        else:
            w = torch.tensor([float(x) for x in items], dtype=torch.float32)
    return w.to(device=device, dtype=torch.float32).contiguous()


def consensus_batched(
    labels: torch.Tensor,
    data_parameters: torch.Tensor,
    *,
    out_dtype: Optional[torch.dtype] = None,
    count_near_ties: bool = False,
    weights_are_softmaxed: bool = False,
    out: Optional[torch.Tensor] = None,
) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
    """Vote for ``T`` targets in one launch.

    labels            ``[T, N, *spatial]`` int64 or uint8, CUDA, the N registered atlas candidates
    data_parameters   ``[T, N]`` fp32 raw data parameters (softmax over N taken in-kernel,
                      consensus.ipynb:103-104) -- or final weights if ``weights_are_softmaxed``
    returns           ``[T, *spatial]`` (dtype of the labels unless ``out_dtype``), and -- if
                      ``count_near_ties`` -- a 0-dim int64 tensor with the number of voxels whose
                      top-two votes differ by < 1e-6 relative (BASELINE.json tolerance clause).
    """
    _b.require_cuda(labels, "labels")
    if labels.dim() < 3:
        raise ValueError(f"labels must be [T, N, *spatial], got {tuple(labels.shape)}")
    T, N = labels.shape[0], labels.shape[1]
    spatial = tuple(labels.shape[2:])
    lab = labels.contiguous()
    dps = data_parameters.detach().to(device=lab.device, dtype=torch.float32).reshape(T, N).contiguous()
    V = lab[0, 0].numel()
    odt = out_dtype or lab.dtype
    if out is None:
        out = torch.empty((T, *spatial), dtype=odt, device=lab.device)
    elif out.dtype != odt or out.numel() != T * V or not out.is_contiguous() or out.device != lab.device:
        raise ValueError("out must be a contiguous [T, *spatial] tensor of the output dtype on the labels' device")
    ties = torch.zeros((), dtype=torch.int64, device=lab.device) if count_near_ties else None
    lib = _b.load()
    with torch.cuda.device(lab.device):
        rc = lib.ds_consensus_vote(
            _b.ptr(lab), _b.label_dtype(lab), _b.ptr(dps), 0 if weights_are_softmaxed else 1, T, N, V,
            _b.ptr(out), _b.label_dtype(out), _b.ptr(ties), _b.stream_ptr(lab.device),
        )
    _b.check(rc, "ds_consensus_vote")
    return (out, ties) if count_near_ties else out


def _flatten_sparse(lbls: Sequence[torch.Tensor], device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, tuple]:
    """Sparse-COO label tensors -> (linear voxel index, values, segment offsets) on ``device``.
    ``coalesce()`` sorts the indices lexicographically, i.e. by linear index."""
    shape = tuple(lbls[0].shape)
    strides = []
    acc = 1
    for s in reversed(shape):
        strides.append(acc)
        acc *= s
    strides = torch.tensor(list(reversed(strides)), dtype=torch.int64, device=device)
    idx_parts, val_parts, counts = [], [], []
    for l in lbls:
        if tuple(l.shape) != shape:
            raise ValueError("all label candidates must have the same shape")
        l = l.coalesce() if not l.is_coalesced() else l
        ind = l.indices().to(device)
        idx_parts.append((ind * strides[:, None]).sum(0))
        val_parts.append(l.values().to(device=device, dtype=torch.int64))
        counts.append(ind.shape[1])
    voxel_index = torch.cat(idx_parts) if idx_parts else torch.empty(0, dtype=torch.int64, device=device)
    values = torch.cat(val_parts)
    offsets = torch.zeros(len(lbls) + 1, dtype=torch.int64)
    offsets[1:] = torch.tensor(counts, dtype=torch.int64).cumsum(0)
    return voxel_index.contiguous(), values.contiguous(), offsets.to(device), shape


def consensus_sparse(
    lbls: Sequence[torch.Tensor],
    weighting_list,
    *,
    device=None,
    out_dtype: torch.dtype = torch.int64,
    count_near_ties: bool = False,
):
    """Vote straight from sparse-COO candidates (the snapshot's ``modified_labels[i]``) without
    ``torch.stack`` + ``to_dense`` (consensus.ipynb:102,105): only the stored elements are moved to
    and read on the GPU.  Same result bits as the dense kernel."""
    dev = _cuda_device(device if device is not None else (lbls[0].device if lbls[0].is_cuda else None))
    voxel_index, values, offsets, shape = _flatten_sparse(lbls, dev)
    w = _weights_tensor(weighting_list, dev)
    N = len(lbls)
    if w.numel() != N:
        raise ValueError(f"{N} label candidates but {w.numel()} weights")
    V = 1
    for s in shape:
        V *= s
    out = torch.empty(shape, dtype=out_dtype, device=dev)
    ties = torch.zeros((), dtype=torch.int64, device=dev) if count_near_ties else None
    lib = _b.load()
    with torch.cuda.device(dev):
        rc = lib.ds_consensus_vote_sparse(
            _b.ptr(voxel_index), _b.ptr(values), _b.ptr(offsets), _b.ptr(w), 1, 1, N, V,
            _b.ptr(out), _b.label_dtype(out), _b.ptr(ties), _b.stream_ptr(dev),
        )
    _b.check(rc, "ds_consensus_vote_sparse")
    return (out, ties) if count_near_ties else out


_host_staging = {}  # one pinned staging buffer (the last size used) + the event of the copy that last read it


def _upload_host_candidates(lbls, dev) -> Optional[torch.Tensor]:
    """Dense int64 candidate volumes in HOST memory (the notebook's ``DEVICE='cpu'``, consensus.ipynb:30, 382-392)
    -> ``[1, N, *S]`` uint8 on ``dev``: each volume is packed to one byte per voxel by ``ds_host_narrow_labels``
    straight into a pinned staging buffer and the stack crosses PCIe in ONE copy of N x V bytes (instead of N copies
    of 8 V bytes plus a device-side ``torch.stack``).  Returns None when the fast path does not apply (device
    inputs, other dtypes, labels above 255): the caller then uploads the volumes as they are."""
    from .host_io import narrow_labels_host

    if any(l.is_cuda or l.dtype != torch.int64 or l.shape != lbls[0].shape for l in lbls):
        return None
    n, v = len(lbls), lbls[0].numel()
    if v == 0:
        return None
    slot = _host_staging.get("slot")
    if slot is None or slot[0].numel() < n * v:
        _host_staging.clear()
        slot = [torch.empty(n * v, dtype=torch.uint8).pin_memory(), None]
        _host_staging["slot"] = slot
    if slot[1] is not None:
        slot[1].synchronize()
    staging = slot[0][: n * v].view(n, v)
    try:
        for i, l in enumerate(lbls):
            narrow_labels_host(l, staging[i])
    except ValueError:
        return None  # a label above 255: not this path
    with torch.cuda.device(dev):
        out = torch.empty((1, n, *lbls[0].shape), dtype=torch.uint8, device=dev)
        out.view(n, v).copy_(staging, non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record(torch.cuda.current_stream(dev))
    return out


def calc_dp_consensus(lbl_list: Sequence[torch.Tensor], weighting_list, *, device=None, return_near_ties: bool = False):
    """Drop-in for the reference's ``calc_dp_consensus`` (consensus.ipynb:100-109): fuse N registered
    label candidates of one target by the softmax of their data parameters and threshold the vote
    at 0.5.  ``lbl_list``: N tensors ``[D,H,W]`` (any rank), int64 or uint8, sparse COO (as stored
    in ``train_label_snapshot.pth``) or dense; ``weighting_list``: N raw data parameters (floats,
    0-dim or ``[1]`` tensors).  Returns a dense int64 tensor on the device of ``lbl_list[0]``."""
    lbls = list(lbl_list)
    if not lbls:
        raise ValueError("lbl_list is empty")
    src_device = lbls[0].device
    if lbls[0].is_sparse:
        res = consensus_sparse(lbls, weighting_list, device=device, count_near_ties=return_near_ties)
    else:
        dev = _cuda_device(device if device is not None else (src_device if src_device.type == "cuda" else None))
        if any(l.dtype != lbls[0].dtype for l in lbls):
            lbls = [l.long() for l in lbls]
        if lbls[0].dtype not in (torch.int64, torch.uint8):
            lbls = [l.long() for l in lbls]
        stack = _upload_host_candidates(lbls, dev)  # int64 volumes in host memory: packed to uint8, one copy
        if stack is None:
            stack = torch.stack([l.to(dev, non_blocking=True) for l in lbls]).unsqueeze(0)  # [1, N, *S]
        w = _weights_tensor(weighting_list, dev)
        if w.numel() != len(lbls):
            raise ValueError(f"{len(lbls)} label candidates but {w.numel()} weights")
        res = consensus_batched(stack, w.reshape(1, -1), out_dtype=torch.int64, count_near_ties=return_near_ties)
        res = (res[0][0], res[1]) if return_near_ties else res[0]
    if return_near_ties:
        return res[0].to(src_device), int(res[1].item())
    return res.to(src_device)


def staple_batched(labels: torch.Tensor, *, max_iterations: int = 200, foreground: int = 1, return_probability: bool = False):
    """Binary STAPLE (EM) for ``T`` targets at once (``ds_staple``): ``labels [T, N, *spatial]`` int64 / uint8 on the
    GPU, ``N <= 64`` raters.  Returns ``(consensus [T, *spatial], sensitivity [T, N] float64, specificity [T, N] float64,
    iterations [T] int32)`` (+ the probability map ``[T, *spatial]`` fp32 with ``return_probability``), all on the
    device.  The comparison baseline of the consensus notebook (consensus.ipynb:112-124: SimpleITK's
    ``STAPLEImageFilter`` with foreground 1.0 and at most 200 iterations) -- PARITY UNPINNED, see oracle/staple.py."""
    _b.require_cuda(labels, "labels")
    if labels.dim() < 3:
        raise ValueError(f"labels must be [T, N, *spatial], got {tuple(labels.shape)}")
    T, N = labels.shape[0], labels.shape[1]
    spatial = tuple(labels.shape[2:])
    V = labels[0, 0].numel()
    stride = (V + 3) // 4 * 4  # the kernel reads whole words: planes padded to a multiple of 4 voxels (padding votes nothing)
    lab = labels.reshape(T, N, V)
    lab = torch.nn.functional.pad(lab, (0, stride - V)) if stride != V else lab.contiguous()
    dev = lab.device
    out = torch.empty((T, stride), dtype=lab.dtype, device=dev)
    prob = torch.empty((T, stride), dtype=torch.float32, device=dev) if return_probability else None
    sens = torch.zeros((T, N), dtype=torch.float64, device=dev)
    spec = torch.zeros((T, N), dtype=torch.float64, device=dev)
    iters = torch.zeros(T, dtype=torch.int32, device=dev)
    lib = _b.load()
    ws = torch.empty((int(lib.ds_staple_workspace_bytes(T, N)) + 7) // 8, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.ds_staple(_b.ptr(lab), _b.label_dtype(lab), T, N, V, stride, int(foreground), int(max_iterations), _b.ptr(out),
                           _b.label_dtype(out), _b.ptr(prob), _b.ptr(sens), _b.ptr(spec), _b.ptr(iters), _b.ptr(ws),
                           ws.numel() * 8, _b.stream_ptr(dev))
    _b.check(rc, "ds_staple")
    res = (out[:, :V].reshape(T, *spatial), sens, spec, iters)
    return res + (prob[:, :V].reshape(T, *spatial),) if return_probability else res


def calc_staple_consensus(lbl_list: Sequence[torch.Tensor], *, max_iterations: int = 200, device=None):
    """Drop-in for the notebook's ``calc_staple_consensus(lbl_list)`` (consensus.ipynb:112-124): returns
    ``(consensus int64 [D,H,W] on the labels' device, sensitivities, specificities)`` -- the two lists of python
    floats that ``GetSensitivity()`` / ``GetSpecificity()`` return there.  ``lbl_list``: N sparse-COO or dense label
    volumes; foreground is label value 1 (``SetForegroundValue(1.0)``), at most ``max_iterations`` EM iterations
    (``SetMaximumIterations(200)``)."""
    lbls = list(lbl_list)
    if not lbls:
        raise ValueError("lbl_list is empty")
    src_device = lbls[0].device
    dev = _cuda_device(device if device is not None else (src_device if src_device.type == "cuda" else None))
    dense = [(l.to(dev).to_dense() if l.is_sparse else l.to(dev)) for l in lbls]
    stack = torch.stack([d if d.dtype == torch.uint8 else d.to(torch.int64) for d in dense]).unsqueeze(0)
    if stack.dtype == torch.int64:  # the notebook casts to int16 anyway (:118); bytes are 8x lighter
        stack = stack.clamp_(0, 255).to(torch.uint8)
    out, sens, spec, _ = staple_batched(stack, max_iterations=max_iterations)
    cons = out[0].long().to(src_device)
    return cons, [float(x) for x in sens[0].cpu()], [float(x) for x in spec[0].cpu()]


def consensus_multiclass(labels: torch.Tensor, data_parameters: torch.Tensor, num_classes: int,
                         *, out_dtype: Optional[torch.dtype] = None, weights_are_softmaxed: bool = False) -> torch.Tensor:
    """EXTENSION without a reference counterpart: per-class weighted vote + argmax (ties -> lowest
    class index) over ``labels [T, N, *spatial]``."""
    _b.require_cuda(labels, "labels")
    T, N = labels.shape[0], labels.shape[1]
    lab = labels.contiguous()
    dps = data_parameters.detach().to(device=lab.device, dtype=torch.float32).reshape(T, N).contiguous()
    out = torch.empty((T, *lab.shape[2:]), dtype=out_dtype or lab.dtype, device=lab.device)
    lib = _b.load()
    with torch.cuda.device(lab.device):
        rc = lib.ds_consensus_vote_multiclass(
            _b.ptr(lab), _b.label_dtype(lab), _b.ptr(dps), 0 if weights_are_softmaxed else 1, T, N,
            lab[0, 0].numel(), int(num_classes), _b.ptr(out), _b.label_dtype(out), _b.stream_ptr(lab.device),
        )
    _b.check(rc, "ds_consensus_vote_multiclass")
    return out


def _sorted_unique(ind: torch.Tensor, shape) -> bool:
    """True if COO indices ``[ndim, nnz]`` are strictly increasing in row-major order, i.e. the tensor
    is coalesced even if its flag says otherwise (``torch.stack`` of coalesced tensors, as written by
    main_deep_staple.py:1025-1027, produces exactly that).  O(nnz), no sort."""
    if ind.shape[1] < 2:
        return True
    lin = torch.zeros(ind.shape[1], dtype=torch.int64, device=ind.device)
    for d, size in enumerate(shape):
        lin = lin * int(size) + ind[d]
    return bool((lin[1:] > lin[:-1]).all())


def _coalesced_parts(st: torch.Tensor):
    """(indices, values) of a sparse COO tensor in coalesced order, sorting only when really needed."""
    if st.is_coalesced():
        return st.indices(), st.values()
    ind, val = st._indices(), st._values()
    if _sorted_unique(ind, st.shape):
        return ind, val
    st = st.coalesce()
    return st.indices(), st.values()


class _SparseRows:
    """``stacked[k]`` for a stacked sparse-COO tensor ``[K, *S]`` in O(nnz of row k): torch's own
    sparse indexing scans all stored elements for every row, which makes the notebook's grouping loop
    (consensus.ipynb:353-377) quadratic in the snapshot size.  Rows come out as coalesced sparse
    tensors with the same indices / values ``stacked[k]`` would give."""

    def __init__(self, stacked: torch.Tensor):
        self.dense = None
        if not (isinstance(stacked, torch.Tensor) and stacked.is_sparse):
            self.dense = stacked
            return
        self.shape = tuple(stacked.shape[1:])
        ind, self.val = _coalesced_parts(stacked)
        counts = torch.bincount(ind[0], minlength=stacked.shape[0])
        self.ends = torch.cumsum(counts, 0).tolist()
        self.rest = ind[1:]

    def __getitem__(self, k: int) -> torch.Tensor:
        if self.dense is not None:
            return self.dense[k]
        lo = self.ends[k - 1] if k > 0 else 0
        hi = self.ends[k]
        return torch.sparse_coo_tensor(self.rest[:, lo:hi], self.val[lo:hi], self.shape, is_coalesced=True,
                                       check_invariants=False)


def group_snapshot(snapshot: Dict) -> Dict[str, Dict]:
    """consensus.ipynb:353-377: group the snapshot by fixed (target) id.  Ids look like
    ``'126l:m102l'``: fixed id ``_id[:4]``, moving (atlas) id ``_id[6:]``; the first entry of a fixed
    id supplies ``expert_label`` and ``prediction``."""
    d_ids = snapshot["d_ids"]
    first_index = {}
    for k, _id in enumerate(d_ids):
        first_index.setdefault(_id, k)  # the notebook uses d_ids.index(_id): first occurrence
    labels, preds = _SparseRows(snapshot["labels"]), _SparseRows(snapshot["train_predictions"])
    mods = _SparseRows(snapshot["modified_labels"])
    consensus_dicts: Dict[str, Dict] = {}
    for _id in d_ids:
        k = first_index[_id]
        f_id, m_id = _id[:4], _id[6:]
        fixed = consensus_dicts.get(f_id)
        if fixed is None:
            fixed = {"expert_label": labels[k], "prediction": preds[k]}
        moving = fixed.get(m_id, {})
        moving["warped_label"] = mods[k]
        moving["data_parameter"] = snapshot["data_parameters"][k]
        fixed[m_id] = moving
        consensus_dicts[f_id] = fixed
    return consensus_dicts


def _batched_snapshot_vote(snapshot, dicts, dev):
    """All targets of a snapshot in ONE vote launch straight from the stacked sparse-COO
    ``modified_labels [K, *S]`` (no per-candidate tensors, no densification): returns
    ``(maps uint8 [T, V] on dev, order of the fixed ids)`` or ``None`` when the targets do not all have
    the same number of candidates (then the caller votes target by target)."""
    mod = snapshot["modified_labels"]
    if not (isinstance(mod, torch.Tensor) and mod.is_sparse):
        return None
    d_ids = list(snapshot["d_ids"])
    first = {}
    for k, _id in enumerate(d_ids):
        first.setdefault(_id, k)
    row_of: Dict[str, Dict[str, int]] = {}  # same insertion / overwrite order as group_snapshot
    for _id in d_ids:
        row_of.setdefault(_id[:4], {})[_id[6:]] = first[_id]
    f_ids = list(dicts.keys())
    rows = [list(row_of[f_id].values()) for f_id in f_ids]  # snapshot rows in (target, candidate) order
    n = len(rows[0])
    if n == 0 or any(len(r) != n for r in rows):
        return None
    T = len(rows)
    shape = tuple(mod.shape[1:])
    V = 1
    for d in shape:
        V *= d
    ind, val = _coalesced_parts(mod.to(dev))  # sortedness check / sort on the GPU
    val = val.to(torch.int64)
    strides = []
    acc = 1
    for d in reversed(shape):
        strides.append(acc)
        acc *= d
    strides = torch.tensor(list(reversed(strides)), dtype=torch.int64, device=dev)
    lin = (ind[1:] * strides[:, None]).sum(0)  # linear voxel index; sorted inside every instance (coalesced)
    K = mod.shape[0]
    counts = torch.bincount(ind[0], minlength=K)
    starts = torch.cumsum(counts, 0) - counts
    order = torch.tensor([r for row in rows for r in row], dtype=torch.int64, device=dev)
    seg_len = counts[order]
    seg_off = torch.zeros(T * n + 1, dtype=torch.int64, device=dev)
    seg_off[1:] = torch.cumsum(seg_len, 0)
    total = int(seg_off[-1])
    # gather the stored elements into (target, candidate) segment order
    src = torch.repeat_interleave(starts[order] - seg_off[:-1], seg_len) + torch.arange(total, device=dev)
    voxel_index = lin[src].contiguous()
    values = val[src].contiguous()
    dps = snapshot["data_parameters"].detach().reshape(-1).to(device=dev, dtype=torch.float32)[order].reshape(T, n).contiguous()
    out = torch.empty((T, V), dtype=torch.uint8, device=dev)
    lib = _b.load()
    with torch.cuda.device(dev):
        rc = lib.ds_consensus_vote_sparse(_b.ptr(voxel_index), _b.ptr(values), _b.ptr(seg_off), _b.ptr(dps), 1, T, n, V,
                                          _b.ptr(out), _b.DS_U8, None, _b.stream_ptr(dev))
    _b.check(rc, "ds_consensus_vote_sparse")
    return out, f_ids, shape


def _add_staple(dicts: Dict[str, Dict], device, with_dice: bool) -> None:
    """The STAPLE half of the driver cell (consensus.ipynb:393-396, 408-418): per target ``staple_consensus`` (sparse),
    ``staple_consensus_oracle_dice`` and per moving atlas ``staple_sensitivity``."""
    from .metrics import dice_from_label_maps

    for f_id, fixed in dicts.items():
        ids = [k for k, e in fixed.items() if isinstance(e, dict)]
        cons, sens, _ = calc_staple_consensus([fixed[k]["warped_label"] for k in ids], device=device)
        for k, s in zip(ids, sens):
            fixed[k]["staple_sensitivity"] = s
        fixed["staple_consensus"] = cons.cpu().to_sparse()
        if with_dice:
            expert = fixed["expert_label"]
            expert = expert.to_dense() if expert.is_sparse else expert
            fixed["staple_consensus_oracle_dice"] = dice_from_label_maps(
                cons.cpu().unsqueeze(0), expert.cpu().unsqueeze(0), 2, nan_for_unlabeled_target=False)


def consensus_from_snapshot(snapshot, *, device=None, with_dice: bool = True, with_staple: bool = False) -> Dict[str, Dict]:
    """The consensus driver cell (consensus.ipynb:353-418); ``with_staple`` adds the STAPLE baseline's keys
    (``staple_consensus``, ``staple_consensus_oracle_dice``, ``<m_id>.staple_sensitivity``) from the GPU EM of
    ``calc_staple_consensus`` instead of SimpleITK's filter (parity unpinned, hence off by default):
    ``snapshot`` is the dict saved as ``train_label_snapshot.pth`` (main_deep_staple.py:963-1037) or
    a path to it.  Per target: vote over all its registered candidates, Dice against the expert label,
    results stored under ``dp_consensus`` (sparse, like the notebook) / ``dp_consensus_oracle_dice``.

    When every target has the same number of candidates (the reference's data: N atlases per target)
    the whole snapshot is voted on in one launch from the stacked sparse labels and the Dice counts of
    all targets come from one more launch; otherwise the targets are processed one by one."""
    if isinstance(snapshot, (str, bytes)) or hasattr(snapshot, "__fspath__"):
        from .snapshot import load_snapshot

        snapshot = load_snapshot(snapshot)
    dicts = group_snapshot(snapshot)
    dev = _cuda_device(device)
    batched = _batched_snapshot_vote(snapshot, dicts, dev)
    if batched is not None:
        maps, f_ids, shape = batched
        T = len(f_ids)
        dice = None
        if with_dice:
            from .loss import label_statistics
            from .metrics import dice_from_stats

            experts = []
            for f_id in f_ids:
                e = dicts[f_id]["expert_label"]
                experts.append((e.to(dev).to_dense() if e.is_sparse else e.to(dev)).reshape(-1).to(torch.uint8))
            stats = label_statistics(maps, torch.stack(experts), 2)
            dice = dice_from_stats(stats, nan_for_unlabeled_target=False)  # CPU [T, 2]
        for i, f_id in enumerate(f_ids):
            sp = maps[i].reshape(shape).to_sparse()
            dicts[f_id]["dp_consensus"] = torch.sparse_coo_tensor(sp.indices(), sp.values().long(), shape, is_coalesced=True,
                                                                check_invariants=False).cpu()
            if dice is not None:
                dicts[f_id]["dp_consensus_oracle_dice"] = dice[i : i + 1]
        if with_staple:
            _add_staple(dicts, device, with_dice)
        return dicts
    for f_id, fixed in dicts.items():
        entries = [e for e in fixed.values() if isinstance(e, dict)]
        lbls: List[torch.Tensor] = [e["warped_label"] for e in entries]
        dps = [e["data_parameter"] for e in entries]
        cons = calc_dp_consensus(lbls, dps, device=device)
        fixed["dp_consensus"] = cons.to_sparse()
        if with_dice:
            from .metrics import dice_from_label_maps

            expert = fixed["expert_label"]
            expert = expert.to_dense() if expert.is_sparse else expert
            fixed["dp_consensus_oracle_dice"] = dice_from_label_maps(
                cons.unsqueeze(0), expert.unsqueeze(0), 2, nan_for_unlabeled_target=False
            )
        dicts[f_id] = fixed
    if with_staple:
        _add_staple(dicts, device, with_dice)
    return dicts
