"""Multi-GPU sharding of the hot path: one process per B200, ``torch.distributed`` (NCCL over
NVLink 5 / NVSwitch) for the only two exchanges the path needs.  The reference is single process /
single device (main_deep_staple.py:35-36); this is a new capability, designed from SURVEY.md section 8(e).

* Consensus (consensus.ipynb:382-392): targets are independent -> contiguous block of targets per
  rank, NO collective while voting; optionally one final all-gather of the label maps and one
  all-reduce of the near-tie counter.
* DP loss (main_deep_staple.py:738-762): voxel sums are per instance, but ``weight.mean()`` (:744)
  and the ``L/B`` term of the gradient couple the whole batch.  Per step:
    1. all-gather the instances' dataset indices (B_global x 8 bytes) -> every rank draws the data
       parameters of the WHOLE batch from its embedding replica, so the instance weights are known
       before the streaming pass and forward + backward still fuse into one sweep;
    2. the fused kernel writes this rank's rows of the [B_global, C, 4] statistics;
    3. all-gather of those rows (B_global x C x 32 bytes);
    4. every rank runs the identical O(B) epilogue -> identical loss and identical sparse gradient
       rows for all B_global data parameters on every rank: the embedding / SparseAdam replicas stay
       bit-identical without any gradient all-reduce.
  The exchange helpers take an explicit process group and work on any backend, so the host logic is
  covered by world_size-2 gloo tests on CPU (tests/test_distributed_cpu.py) with the kernel call
  replaced by a stub; on the GPU the kernel calls are the C-ABI ones.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import _binding as _b

__all__ = [
    "FusedExchange",
    "VoteGather",
    "shard_range",
    "gather_batch_indices",
    "exchange_instance_stats",
    "sharded_consensus",
    "dp_weighted_loss_distributed",
    "global_class_counts",
]


class FusedExchange:
    """Peer-mapped exchange buffers for ``ds_dp_loss_fwd_bwd_exchange``: the statistics exchange of the
    data-parallel loss happens INSIDE the loss kernel (stores into every rank's buffer over NVLink /
    NVSwitch peer memory + arrival flags), replacing the NCCL all-gather and the separate epilogue launch.

    One symmetric allocation per (group, B_global, C), made with ``torch.distributed._symmetric_memory``
    (CUDA VMM / fabric handles exchanged through the process group's store).  All ranks must construct
    it collectively and then make the same sequence of loss calls (the step counter ``epoch`` must agree)."""

    _cache = {}

    def __init__(self, b_total: int, num_classes: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        lib = _b.load()
        nbytes = int(lib.ds_dp_loss_exchange_bytes(b_total, num_classes, self.world))
        self.buf = symm_mem.empty((nbytes + 7) // 8, dtype=torch.int64, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        ptrs = [int(x) for x in self.handle.buffer_ptrs]
        offset = self.buf.data_ptr() - ptrs[self.rank]  # the tensor may sit inside a larger symmetric block
        self.peer_bases = torch.tensor([x + offset for x in ptrs], dtype=torch.int64, device=device)
        self.epoch = 0
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every buffer is zeroed before anyone's first flag can arrive

    @classmethod
    def get(cls, b_total: int, num_classes: int, device: torch.device, group=None) -> "FusedExchange":
        """The exchange of (group, global batch, classes), built on first use -- a COLLECTIVE (rendezvous + barrier):
        call it explicitly on every rank before the first step if that step is going to be captured in a CUDA graph.
        At most 8 exchanges are kept (a training run uses one or two); ``clear_cache()`` drops them all."""
        key = (id(group), b_total, num_classes, device.index)
        if key not in cls._cache:
            if len(cls._cache) >= 8:
                cls._cache.pop(next(iter(cls._cache)))
            cls._cache[key] = cls(b_total, num_classes, device, group)
        return cls._cache[key]

    @classmethod
    def clear_cache(cls) -> None:
        cls._cache.clear()

    def check(self) -> None:
        """Raise if this rank ever lost a step (a peer's statistics did not arrive within
        ``DS_EXCHANGE_TIMEOUT_S`` seconds, default 1800): the kernel then makes every output of that and all
        later steps NaN -- GradScaler skips them -- because this replica may no longer have applied the same
        updates as its peers.  One device->host read; call it wherever the training loop already synchronises
        (``dp_loss.item()``, main_deep_staple.py:775) or once per epoch."""
        lost_at = int(self.buf[-1].item())
        if lost_at:
            raise RuntimeError(
                f"fused statistics exchange: rank {self.rank} lost step {lost_at} (a peer did not arrive in time); "
                "all outputs since then are NaN.  Restore the replicas from a checkpoint and call FusedExchange.reset()")

    def reset(self, group=None) -> None:
        """Collective: clear the exchange buffers (step counter, slots, error word) on every rank."""
        torch.cuda.synchronize(self.buf.device)
        dist.barrier(group)
        self.buf.zero_()
        torch.cuda.synchronize(self.buf.device)
        dist.barrier(group)

    def next_epoch(self) -> int:
        """Step counter argument of ``ds_dp_loss_fwd_bwd_exchange``: 0 = the kernel keeps the counter in the
        exchange buffer itself (no host state, so the step can be captured in a CUDA graph)."""
        return 0


class VoteGather:
    """Symmetric ``[T_global, V]`` label-map buffers for ``ds_consensus_vote_gather``: one buffer per rank, every
    buffer mapped into every rank (``torch.distributed._symmetric_memory``: CUDA VMM / fabric handles), so the vote
    kernel stores each consensus map straight into all ranks' copies over NVLink / NVSwitch while it is still reading
    label planes -- the all-gather of the target-sharded consensus without a collective, a padding copy or a ``cat``.

    Construction is COLLECTIVE (rendezvous).  ``barrier()`` orders the ranks on the current stream (signal pads of
    the symmetric allocation; no NCCL kernel)."""

    _cache = {}

    def __init__(self, n_targets_global: int, voxels: int, dtype: torch.dtype, device: torch.device, group=None):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        group = group if group is not None else dist.group.WORLD
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 16:
            raise ValueError("VoteGather supports at most 16 ranks (DS_MAX_PEERS)")
        self.buf = symm_mem.empty((n_targets_global, voxels), dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group)
        ptrs = [int(x) for x in self.handle.buffer_ptrs]
        offset = self.buf.data_ptr() - ptrs[self.rank]  # the tensor may sit inside a larger symmetric block
        self.table = (ctypes.c_void_p * self.world)(*[x + offset for x in ptrs])  # host table the C ABI takes
        torch.cuda.synchronize(device)
        dist.barrier(group)

    @classmethod
    def get(cls, n_targets_global: int, voxels: int, dtype: torch.dtype, device: torch.device, group=None) -> "VoteGather":
        """The buffers of (group, targets, voxels, dtype), built on first use (collective).  At most 4 are kept."""
        key = (id(group), n_targets_global, voxels, dtype, device.index)
        if key not in cls._cache:
            if len(cls._cache) >= 4:
                cls._cache.pop(next(iter(cls._cache)))
            cls._cache[key] = cls(n_targets_global, voxels, dtype, device, group)
        return cls._cache[key]

    @classmethod
    def clear_cache(cls) -> None:
        cls._cache.clear()

    def barrier(self) -> None:
        self.handle.barrier(channel=0)


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of ``n_items`` owned by ``rank``: sizes differ by at most one, the first
    ``n_items % world_size`` ranks get the larger blocks (100 targets over 8 GPUs -> 13,13,13,13,12,...)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _world(group) -> Tuple[int, int]:
    return dist.get_rank(group), dist.get_world_size(group)


def gather_batch_indices(idx_local: torch.Tensor, group=None) -> Tuple[torch.Tensor, int]:
    """All-gather ``b_idxs_dataset`` of every rank -> (``[B_global]`` indices in rank order, row offset
    of this rank).  Every rank must hold the same number of instances."""
    rank, world = _world(group)
    idx_local = idx_local.reshape(-1).contiguous()
    out = torch.empty(world * idx_local.numel(), dtype=idx_local.dtype, device=idx_local.device)
    dist.all_gather_into_tensor(out, idx_local, group=group)
    return out, rank * idx_local.numel()


def exchange_instance_stats(stats_full: torch.Tensor, b_offset: int, b_local: int, group=None) -> torch.Tensor:
    """In-place all-gather of the per-instance statistics: ``stats_full`` is ``[B_global, C, 4]`` with
    this rank's rows ``[b_offset, b_offset + b_local)`` filled; afterwards all rows are filled on
    every rank."""
    mine = stats_full[b_offset : b_offset + b_local]
    dist.all_gather_into_tensor(stats_full, mine.contiguous(), group=group)
    return stats_full


def global_class_counts(local_counts: torch.Tensor, group=None) -> torch.Tensor:
    """Label histogram of the GLOBAL batch (normaliser of the class-weighted CE, :716) from the local one."""
    counts = local_counts.clone()
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def sharded_consensus(
    labels_local: torch.Tensor,
    data_parameters_local: torch.Tensor,
    n_targets_global: int,
    *,
    gather: bool = True,
    count_near_ties: bool = False,
    group=None,
    vote_fn: Optional[Callable] = None,
    fused: Optional[bool] = None,
    out_dtype: Optional[torch.dtype] = None,
):
    """Vote on this rank's block of targets (``labels_local [T_r, N, *S]`` = targets
    ``shard_range(n_targets_global, rank, world)``); with ``gather`` every rank receives the full
    ``[T_global, *S]`` stack.  Returns ``(maps, near_ties)``.

    ``fused`` (default: on for CUDA labels on more than one rank): the gather happens INSIDE the vote kernel --
    ``ds_consensus_vote_gather`` stores every map into all ranks' ``VoteGather`` buffers over NVLink peer memory,
    two stream-ordered barriers around it, no NCCL call; the returned stack is that buffer (valid until the next
    fused call with the same shapes).  ``fused=False``: vote, then one NCCL all-gather.
    ``out_dtype``: dtype of the gathered maps (default: the labels' dtype; ``torch.uint8`` skips the widening pass
    -- the binary vote only ever emits 0 / 1 and the maps always TRAVEL as uint8)."""
    rank, world = _world(group)
    lo, hi = shard_range(n_targets_global, rank, world)
    if labels_local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} owns targets [{lo},{hi}) but got {labels_local.shape[0]} label stacks")
    ties = None
    final_dtype = out_dtype or labels_local.dtype
    if fused is None:
        fused = gather and vote_fn is None and labels_local.is_cuda and world > 1
    if fused:
        if not gather or vote_fn is not None:
            raise ValueError("fused=True is the gather inside the vote kernel: it needs gather=True and no vote_fn")
        return _fused_vote_gather(labels_local, data_parameters_local, n_targets_global, lo, hi, count_near_ties, group,
                                  final_dtype)
    # The binary vote only ever emits 0 / 1: with the gather ahead, int64 label stacks are voted into uint8 maps,
    # which travel 8x lighter through the all-gather (the collective, not the vote, is the expensive part: 1.7 GB
    # per rank for cfg3 as int64) and are widened back to the labels' dtype on arrival.
    compact = gather and vote_fn is None and labels_local.dtype == torch.int64
    if vote_fn is None:
        from .consensus import consensus_batched

        vote_fn = consensus_batched
    if hi > lo:
        kw = {"out_dtype": torch.uint8} if compact else {}
        res = vote_fn(labels_local, data_parameters_local, count_near_ties=count_near_ties, **kw)
        local, ties = res if count_near_ties else (res, None)
    else:
        local = torch.empty((0, *labels_local.shape[2:]), dtype=torch.uint8 if compact else labels_local.dtype,
                            device=labels_local.device)
    if count_near_ties:
        ties = ties if ties is not None else torch.zeros((), dtype=torch.int64, device=labels_local.device)
        ties = ties.clone()
        dist.all_reduce(ties, op=dist.ReduceOp.SUM, group=group)
    if not gather:
        return local, ties
    # ragged blocks (n_targets % world != 0): pad every block to the largest, gather, cut
    per = -(-n_targets_global // world)
    padded = local.contiguous()
    if local.shape[0] != per:
        padded = torch.zeros((per, *local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    full = torch.empty((world * per, *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, padded, group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(n_targets_global, r, world)
        parts.append(full[r * per : r * per + (b - a)])
    full = torch.cat(parts) if len(parts) > 1 and n_targets_global % world else full[:n_targets_global]
    return (full.to(final_dtype) if full.dtype != final_dtype else full), ties


def _fused_vote_gather(labels_local, dps_local, n_targets_global, lo, hi, count_near_ties, group, final_dtype):
    """``sharded_consensus`` with the gather inside the vote kernel (see ``VoteGather``)."""
    dev = labels_local.device
    T_local, N = hi - lo, labels_local.shape[1]
    spatial = tuple(labels_local.shape[2:])
    V = 1
    for s in spatial:
        V *= s
    wire = torch.uint8 if final_dtype in (torch.uint8, torch.int64) else final_dtype
    vg = VoteGather.get(n_targets_global, V, wire, dev, group)
    ties = torch.zeros((), dtype=torch.int64, device=dev) if count_near_ties else None
    lab = labels_local.contiguous()
    dps = dps_local.detach().to(device=dev, dtype=torch.float32).reshape(T_local, N).contiguous() if T_local else None
    lib = _b.load()
    with torch.cuda.device(dev):
        vg.barrier()  # every rank is done reading the previous result out of its buffer
        if T_local:
            rc = lib.ds_consensus_vote_gather(
                _b.ptr(lab), _b.label_dtype(lab), _b.ptr(dps), 1, T_local, N, V, lo, n_targets_global, vg.table, vg.world,
                _b.label_dtype(vg.buf), _b.ptr(ties), _b.stream_ptr(dev))
            _b.check(rc, "ds_consensus_vote_gather")
        vg.barrier()  # every rank's rows have landed in every buffer
    if count_near_ties:
        dist.all_reduce(ties, op=dist.ReduceOp.SUM, group=group)
    full = vg.buf.view(n_targets_global, *spatial)
    return (full.to(final_dtype) if full.dtype != final_dtype else full), ties


class _DistributedDPLoss(torch.autograd.Function):
    """Fused local pass + statistics exchange + replicated epilogue (see module docstring)."""

    @staticmethod
    def forward(ctx, dp_logits, bare_global, target, fixed_global, use_risk, scale, class_w, b_offset, group, fused):
        from .loss import _flat_inputs, _workspace

        z, t, B_local, C, V = _flat_inputs(dp_logits, target)
        dev = z.device
        bare = bare_global.detach().to(torch.float32).reshape(-1).contiguous()
        B_total = bare.numel()
        fixed = None if fixed_global is None else fixed_global.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        cw = None if class_w is None else class_w.detach().to(device=dev, dtype=torch.float32).contiguous()
        need_dlogits = ctx.needs_input_grad[0]
        if need_dlogits and scale is None and z.dtype is not torch.float32:
            raise ValueError("half-precision dp_logits that require grad need grad_scaler=...: their gradient is written by "
                             "the fused pass already scaled (unscaled it would underflow fp16)")
        dlogits = torch.empty_like(z) if need_dlogits else None
        stats = torch.zeros((B_total, C, _b.DS_STATS), dtype=torch.float64, device=dev)
        loss_out = torch.empty(2, dtype=torch.float32, device=dev)
        d_bare = torch.empty(B_total, dtype=torch.float32, device=dev)
        per_inst = torch.empty((B_total, _b.DS_INST), dtype=torch.float32, device=dev)
        lib = _b.load()
        with torch.cuda.device(dev):
            ws = _workspace(dev, B_local, C)
            if fused and B_local != B_total:
                # statistics exchange + epilogue inside the loss kernel (peer memory over NVLink)
                ex = FusedExchange.get(B_total, C, dev, group)
                rc = lib.ds_dp_loss_fwd_bwd_exchange(
                    _b.ptr(z), _b.logit_dtype(z), _b.ptr(t), _b.label_dtype(t), None, B_local, C, V, _b.ptr(bare),
                    _b.ptr(fixed), B_total, b_offset, _b.ptr(cw), 1 if use_risk else 0, _b.ptr(scale),
                    _b.ptr(dlogits), _b.ptr(stats), _b.ptr(loss_out), _b.ptr(d_bare), _b.ptr(per_inst),
                    _b.ptr(ex.peer_bases), ex.rank, ex.world, ex.next_epoch(), _b.ptr(ws), ws.numel() * 8,
                    _b.stream_ptr(dev),
                )
                _b.check(rc, "ds_dp_loss_fwd_bwd_exchange")
            else:
                rc = lib.ds_dp_loss_fwd_bwd(
                    _b.ptr(z), _b.logit_dtype(z), _b.ptr(t), _b.label_dtype(t), None, B_local, C, V, _b.ptr(bare),
                    _b.ptr(fixed), B_total, b_offset, _b.ptr(cw), 1 if use_risk else 0, _b.ptr(scale),
                    _b.ptr(dlogits), _b.ptr(stats), _b.ptr(loss_out), _b.ptr(d_bare), _b.ptr(per_inst), _b.ptr(ws),
                    ws.numel() * 8, _b.stream_ptr(dev),
                )
                _b.check(rc, "ds_dp_loss_fwd_bwd")
                if B_local != B_total:
                    exchange_instance_stats(stats, b_offset, B_local, group)
                    rc = lib.ds_dp_loss_finalize(
                        _b.ptr(stats), B_total, C, V, _b.ptr(bare), _b.ptr(fixed), _b.ptr(cw), 1 if use_risk else 0,
                        _b.ptr(scale), _b.ptr(loss_out), _b.ptr(d_bare), _b.ptr(per_inst), _b.stream_ptr(dev),
                    )
                    _b.check(rc, "ds_dp_loss_finalize")
        ctx.assumed = scale          # scale inside d_bare (never rewritten)
        ctx.dlogits_scale = scale    # scale inside the stored dlogits (rewritten by a reconciling backward)
        ctx.bare_shape = bare_global.shape
        ctx.logits_shape = dp_logits.shape
        ctx.save_for_backward(d_bare, dlogits)
        ce = loss_out[1]
        ctx.mark_non_differentiable(ce, per_inst, stats)
        return loss_out[0], ce, per_inst, stats

    @staticmethod
    def backward(ctx, grad_loss, _a, _b2, _c):
        d_bare, dlogits = ctx.saved_tensors
        dev = d_bare.device
        g = grad_loss
        if g.dtype is not torch.float32 or g.dim() != 0 or g.requires_grad:
            g = g.detach().to(torch.float32).reshape(())
        grad_bare = torch.empty_like(d_bare)
        lib = _b.load()
        with torch.cuda.device(dev):
            fused = dlogits
            if dlogits is not None and ctx.dlogits_scale is not ctx.assumed:  # a second backward after a reconciling one
                one = ctx.dlogits_scale if ctx.dlogits_scale is not None else torch.ones((), dtype=torch.float32, device=dev)
                rc = lib.ds_rescale_inplace(_b.ptr(dlogits), _b.logit_dtype(dlogits), dlogits.numel(), _b.ptr(g), _b.ptr(one),
                                            _b.stream_ptr(dev))
                _b.check(rc, "ds_rescale_inplace")
                fused = None
            # one launch: d_bare x (g / assumed) into a fresh tensor, and the stored dlogits rescaled only if g differs
            rc = lib.ds_dp_loss_backward_reconcile(
                _b.ptr(fused), _b.logit_dtype(fused) if fused is not None else _b.DS_F32,
                fused.numel() if fused is not None else 0, _b.ptr(d_bare), _b.ptr(grad_bare), d_bare.numel(), _b.ptr(g),
                _b.ptr(ctx.assumed), _b.stream_ptr(dev))
            _b.check(rc, "ds_dp_loss_backward_reconcile")
        grad_logits = None
        if dlogits is not None:
            ctx.dlogits_scale = g
            grad_logits = dlogits.view(ctx.logits_shape)
        return grad_logits, grad_bare.reshape(ctx.bare_shape), None, None, None, None, None, None, None, None


def dp_weighted_loss_distributed(
    dp_logits: torch.Tensor,
    b_seg_modified: torch.Tensor,
    embedding: torch.nn.Embedding,
    b_idxs_dataset: torch.Tensor,
    fixed_weighting: Optional[torch.Tensor] = None,
    use_risk_regularization: bool = True,
    *,
    grad_scaler=None,
    class_weights: Optional[torch.Tensor] = None,
    group=None,
    return_all: bool = False,
    exchange: str = "fused",
):
    """Data-parallel form of ``dp_weighted_loss``: each rank passes ITS instances (``dp_logits``,
    ``b_seg_modified``, ``b_idxs_dataset``), the replicated ``embedding`` and the FULL
    ``fixed_weighting`` table (:653).  Returns the loss of the GLOBAL batch (identical on every rank);
    ``backward`` delivers sparse gradient rows for all ``B_global`` data parameters on every rank (so
    the replicas stay identical) and, if ``dp_logits`` requires grad, this rank's logits gradient.
    ``exchange``: ``"fused"`` -- the statistics travel inside the loss kernel over NVLink peer memory
    (``FusedExchange``); ``"nccl"`` -- fused kernel + NCCL all-gather + epilogue kernel."""
    if exchange not in ("fused", "nccl"):
        raise ValueError("exchange must be 'fused' or 'nccl'")
    from .loss import DPLossOutput, _scale_tensor

    idx_global, b_offset = gather_batch_indices(b_idxs_dataset.to(dp_logits.device), group)
    bare_global = embedding(idx_global).squeeze()  # :741 on the global batch
    fixed_global = fixed_weighting[idx_global] if fixed_weighting is not None else None  # :748
    scale = _scale_tensor(grad_scaler, dp_logits.device)
    loss, ce, per_inst, stats = _DistributedDPLoss.apply(
        dp_logits, bare_global, b_seg_modified, fixed_global, bool(use_risk_regularization), scale, class_weights,
        b_offset, group, exchange == "fused",
    )
    if not return_all:
        return loss
    return DPLossOutput(
        dp_loss=loss, ce_loss=ce, per_instance_ce=per_inst[:, _b.DS_INST_CE], fg_pred=per_inst[:, _b.DS_INST_FG],
        instance_weight=per_inst[:, _b.DS_INST_WEIGHT], inst_stats=stats,
    )
