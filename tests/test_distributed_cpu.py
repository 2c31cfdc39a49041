"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in deep_staple_b200/distributed.py:
sharding, the index / statistics exchanges and the ragged consensus gather.  The CUDA kernels
are replaced by oracle-based stand-ins here; the same functions run over NCCL on the GPU box."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)


def test_shard_range_partitions_contiguously():
    from deep_staple_b200.distributed import shard_range

    for n, w in ((100, 8), (13, 2), (5, 8), (0, 3), (64, 4)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert [shard_range(100, r, 8)[1] - shard_range(100, r, 8)[0] for r in range(8)] == [13, 13, 13, 13, 12, 12, 12, 12]
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)


def _consensus_worker(rank, world, port, T, out_dir):
    import deep_staple_b200.distributed as dd
    from oracle import consensus as oc

    _init(rank, world, port)
    gen = torch.Generator().manual_seed(0)
    N, shape = 6, (4, 4, 8)
    labels = (torch.rand((T, N, *shape), generator=gen) < 0.4).long()
    dps = torch.randn((T, N), generator=gen)
    lo, hi = dd.shard_range(T, rank, world)

    def vote_fn(lab, dp, count_near_ties=False):  # stand-in for consensus_batched
        res = torch.stack([oc.calc_dp_consensus(list(lab[t]), [float(x) for x in dp[t]]) for t in range(lab.shape[0])])
        ties = torch.tensor(sum(int(oc.near_tie_mask(list(lab[t]), [float(x) for x in dp[t]]).sum()) for t in range(lab.shape[0])))
        return (res, ties) if count_near_ties else res

    full, ties = dd.sharded_consensus(labels[lo:hi], dps[lo:hi], T, gather=True, count_near_ties=True, vote_fn=vote_fn)
    ref = vote_fn(labels, dps)
    assert full.shape == ref.shape and torch.equal(full, ref), f"rank {rank}: gathered consensus differs"
    local, _ = dd.sharded_consensus(labels[lo:hi], dps[lo:hi], T, gather=False, vote_fn=vote_fn)
    assert torch.equal(local, ref[lo:hi])
    total_ties = sum(int(oc.near_tie_mask(list(labels[t]), [float(x) for x in dps[t]]).sum()) for t in range(T))
    assert int(ties) == total_ties
    open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [5, 4, 1])
def test_sharded_consensus_gloo_world2(tmp_path, T):
    port = _free_port()
    mp.spawn(_consensus_worker, args=(2, port, T, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def _loss_worker(rank, world, port, out_dir):
    import deep_staple_b200.distributed as dd
    from oracle import dp_loss as ol

    _init(rank, world, port)
    gen = torch.Generator().manual_seed(1)
    B_local, C, shape, M = 2, 3, (4, 4, 4), 12
    B = B_local * world
    logits = 2 * torch.randn((B, C, *shape), generator=gen)
    target = torch.randint(0, C, (B, *shape), generator=gen)
    emb_w = torch.randn((M, 1), generator=gen)
    idx = torch.randperm(M, generator=gen)[:B]
    fixed = torch.rand(M, generator=gen) * 3 + 1
    V = math.prod(shape)
    sl = slice(rank * B_local, (rank + 1) * B_local)

    # step 1: every rank learns the dataset indices of the GLOBAL batch
    idx_global, b_offset = dd.gather_batch_indices(idx[sl])
    assert torch.equal(idx_global, idx) and b_offset == rank * B_local

    # step 2 (kernel stand-in): this rank's rows of the [B, C, 4] statistics
    stats = torch.zeros((B, C, 4), dtype=torch.float64)
    z, t = logits[sl].double(), target[sl]
    nll = torch.nn.functional.cross_entropy(z, t, reduction="none")
    pred = z.argmax(1)
    for b in range(B_local):
        for c in range(C):
            stats[b_offset + b, c, 0] = nll[b][t[b] == c].sum()
            stats[b_offset + b, c, 1] = (t[b] == c).sum()
            stats[b_offset + b, c, 2] = (pred[b] == c).sum()
            stats[b_offset + b, c, 3] = ((pred[b] == c) & (t[b] == c)).sum()

    # step 3: in-place all-gather of the rows
    dd.exchange_instance_stats(stats, b_offset, B_local)
    assert float(stats[:, :, 1].sum()) == B * V  # every voxel of every instance counted exactly once

    # step 4: replicated epilogue (closed form) == the reference's loss on the global batch
    per_inst = stats[:, :, 0].sum(1) / V
    fg = stats[:, 1:, 2].sum(1)
    loss, grad, _ = ol.closed_form(per_inst, fg, emb_w.reshape(-1)[idx_global], fixed[idx_global], V, True, 4.0)
    ref = ol.run_loss(logits, target, emb_w, idx, fixed, None, True, 4.0, False)
    assert abs(float(loss) - float(ref.dp_loss)) <= 1e-5 * abs(float(ref.dp_loss))
    assert torch.allclose(grad.float(), ref.d_bare_weight, rtol=1e-4, atol=1e-7)
    # global label histogram for the class-weighted CE normaliser
    counts = dd.global_class_counts(stats[sl, :, 1].sum(0))
    assert torch.equal(counts, torch.bincount(target.reshape(-1), minlength=C).double())
    open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    dist.destroy_process_group()


def test_distributed_loss_exchange_gloo_world2(tmp_path):
    port = _free_port()
    mp.spawn(_loss_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
