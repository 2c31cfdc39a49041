"""Worker of tests/test_gpu_distributed.py (launched by torch.distributed.run, one rank per GPU, NCCL):
the data-parallel loss and the target-sharded consensus must reproduce the single-GPU results."""
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

import deep_staple_b200 as ds  # noqa: E402
from deep_staple_b200 import distributed as dd  # noqa: E402


def full_size_case(name, rank, world, dev, B_local, C, shape, dtype=torch.float32, need_dlogits=True):
    """A BASELINE-sized data-parallel step against the SINGLE-PROCESS CPU ORACLE (oracle.dp_loss.run_loss =
    main_deep_staple.py:716,738-762 on the gathered global batch), not only against our own 1-GPU kernel:
    rank 0 generates the global batch and runs the oracle, every rank gets its instances, runs the fused-exchange
    step and checks loss, ALL global data-parameter gradient rows and its own logits gradient.
    Tolerances: BASELINE.json north_star (loss 1e-5, gradients 1e-4 relative)."""
    from oracle import dp_loss as ol

    B = B_local * world
    V = shape[0] * shape[1] * shape[2]
    M, g_scale = B + 5, 4096.0
    gen = torch.Generator().manual_seed(100 + C)
    emb_w = torch.randn((M, 1), generator=gen)
    idx = torch.randperm(M, generator=gen)[:B]
    fixed = torch.rand(M, generator=gen) * 3 + 1
    cw = torch.rand(C, generator=gen) + 0.5
    z_local = torch.empty((B_local, C, *shape), dtype=dtype, device=dev)
    t_local = torch.empty((B_local, *shape), dtype=torch.int64, device=dev)
    ref_dz = torch.empty((B_local, C, *shape), dtype=dtype, device=dev) if need_dlogits else None
    scalars = torch.empty(2 + 3 * B, dtype=torch.float64, device=dev)  # dp_loss, ce_loss, d_bare[B], per_inst[B], fg[B]
    if rank == 0:
        probs = torch.tensor([0.94, 0.05, 0.01][:C]) if C == 3 else torch.tensor([0.95, 0.05])
        logits = (2 * torch.randn((B, C, *shape), generator=gen)).to(dtype)
        target = torch.multinomial(probs, B * V, replacement=True, generator=gen).reshape(B, *shape)
        torch.set_num_threads(os.cpu_count() or 1)
        ref = ol.run_loss(logits, target, emb_w, idx, fixed, cw, True, g_scale, need_dlogits)
        scalars.copy_(torch.cat([ref.dp_loss.reshape(1).double(), ref.ce_loss.reshape(1).double(), ref.d_bare_weight.double(),
                                 ref.per_inst_ce.double(), ref.fg_pred.double()]))
        z_parts = [logits[r * B_local:(r + 1) * B_local].contiguous().to(dev) for r in range(world)]
        t_parts = [target[r * B_local:(r + 1) * B_local].contiguous().to(dev) for r in range(world)]
        g_parts = [ref.d_logits[r * B_local:(r + 1) * B_local].contiguous().to(dev) for r in range(world)] if need_dlogits else None
    else:
        z_parts = t_parts = g_parts = None
    dist.scatter(z_local, z_parts, src=0)
    dist.scatter(t_local, t_parts, src=0)
    if need_dlogits:
        dist.scatter(ref_dz, g_parts, src=0)
    dist.broadcast(scalars, src=0)
    ref_loss, ref_ce = scalars[0], scalars[1]
    ref_dbare, ref_pi, ref_fg = scalars[2:2 + B], scalars[2 + B:2 + 2 * B], scalars[2 + 2 * B:]

    def rel(a, b):
        a, b = a.double().reshape(-1), b.double().reshape(-1)
        return float((a - b).abs().max() / (b.abs().max() + 1e-300))

    sl = slice(rank * B_local, (rank + 1) * B_local)
    for exchange in ("fused", "nccl"):
        emb = nn.Embedding(M, 1, sparse=True, _weight=emb_w.clone()).to(dev)
        z = z_local.clone().requires_grad_(need_dlogits)
        out = dd.dp_weighted_loss_distributed(z, t_local, emb, idx[sl].to(dev), fixed.to(dev), True, grad_scaler=g_scale,
                                              class_weights=cw.to(dev), return_all=True, exchange=exchange)
        (out.dp_loss * g_scale).backward()
        rows = emb.weight.grad.coalesce().to_dense().reshape(-1)[idx.to(dev)]
        errs = {"dp_loss": rel(out.dp_loss, ref_loss), "ce_loss": rel(out.ce_loss, ref_ce), "d_bare": rel(rows, ref_dbare),
                "per_inst_ce": rel(out.per_instance_ce, ref_pi)}
        fg_equal = bool(torch.equal(out.fg_pred.double(), ref_fg))
        if need_dlogits:
            errs["d_logits"] = rel(z.grad.float(), ref_dz.float())
            # per-instance-normalised: every instance's gradient against ITS OWN magnitude (weights differ by instance)
            per = (z.grad.float() - ref_dz.float()).abs().flatten(1).amax(1) / ref_dz.float().abs().flatten(1).amax(1)
            errs["d_logits_per_instance"] = float(per.max())
        tol_g = 1e-4 if dtype == torch.float32 else 2e-3
        ok = (errs["dp_loss"] < 1e-5 and errs["ce_loss"] < 1e-5 and errs["d_bare"] < 1e-4 and errs["per_inst_ce"] < 1e-5
              and fg_equal and errs.get("d_logits", 0.0) < tol_g and errs.get("d_logits_per_instance", 0.0) < tol_g)
        print(f"[{name}] world {world} rank {rank} B_local {B_local} C {C} shape {tuple(shape)} {str(dtype)[6:]} "
              f"exchange {exchange}: " + " ".join(f"{k} {v:.2e}" for k, v in errs.items()) + f" fg_equal {fg_equal} "
              f"-> {'OK' if ok else 'FAIL'}", flush=True)
        assert ok, (name, exchange, errs, fg_equal)
    dist.barrier()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cases = [a for a in sys.argv[1:] if not a.startswith("-")] or ["small"]
    if "cfg5" in cases:
        # BASELINE configs[4] parity variant: 3 classes, 192x160x160, ONE instance per GPU (B_global = world), in-line
        full_size_case("cfg5-parity", rank, world, dev, 1, 3, (192, 160, 160))
        full_size_case("cfg5-parity-fp16-ool", rank, world, dev, 1, 3, (192, 160, 160), torch.float16, need_dlogits=False)
    if "cfg2" in cases:
        # BASELINE configs[1] per GPU: 8 instances x 2 classes x 128^3
        full_size_case("cfg2-dp", rank, world, dev, 8, 2, (128, 128, 128))
    if "small" not in cases:
        if rank == 0:
            print("DIST_OK", world)
        dist.destroy_process_group()
        return

    # ---- data-parallel loss (cfg5 shape class: 3 classes) -------------------------------------------
    gen = torch.Generator().manual_seed(0)
    B_local, C, shape, M = 2, 3, (16, 20, 24), 40
    B = B_local * world
    logits = (2 * torch.randn((B, C, *shape), generator=gen)).to(dev)
    target = torch.randint(0, C, (B, *shape), generator=gen).to(dev)
    emb_w = torch.randn((M, 1), generator=gen)
    idx = torch.randperm(M, generator=gen)[:B].to(dev)
    fixed = (torch.rand(M, generator=gen) * 3 + 1).to(dev)
    cw = (torch.rand(C, generator=gen) + 0.5).to(dev)
    sl = slice(rank * B_local, (rank + 1) * B_local)

    emb1 = nn.Embedding(M, 1, sparse=True, _weight=emb_w.clone()).to(dev)
    z1 = logits.clone().requires_grad_(True)
    ref = ds.dp_weighted_loss(z1, target, (emb1, idx), fixed[idx], True, grad_scaler=128.0, class_weights=cw,
                              return_all=True)
    (ref.dp_loss * 128.0).backward()

    def close(a, b, rtol, what):
        err = float((a.double() - b.double()).abs().max())
        scale = float(b.double().abs().max())
        assert err <= rtol * scale + 1e-12, f"rank {rank} {what}: err {err:.3e} scale {scale:.3e}"

    # "fused": statistics exchange inside the loss kernel over NVLink peer memory (several steps, so the
    # double-buffered slots and the epoch flags are exercised); "nccl": all-gather + epilogue launch
    for exchange, repeats in (("fused", 5), ("nccl", 1)):
        for it in range(repeats):
            emb = nn.Embedding(M, 1, sparse=True, _weight=emb_w.clone()).to(dev)
            z = logits[sl].clone().requires_grad_(True)
            out = dd.dp_weighted_loss_distributed(z, target[sl], emb, idx[sl], fixed, True, grad_scaler=128.0,
                                                  class_weights=cw, return_all=True, exchange=exchange)
            (out.dp_loss * 128.0).backward()
            what = f"{exchange}#{it}"
            close(out.dp_loss, ref.dp_loss, 1e-6, what + " dp_loss")
            close(out.ce_loss, ref.ce_loss, 1e-6, what + " ce_loss")
            assert torch.equal(out.inst_stats, ref.inst_stats), what + ": statistics must be bit-identical (exact integer reduction)"
            g, g1 = emb.weight.grad.coalesce(), emb1.weight.grad.coalesce()
            assert g.is_sparse and torch.equal(g.indices(), g1.indices())  # ALL global rows on every rank
            close(g.values(), g1.values(), 1e-6, what + " sparse data-parameter gradient")
            close(z.grad, z1.grad[sl], 1e-6, what + " local logits gradient")
    # ---- the fused-exchange step has no host-side state (epoch 0 = counter in the exchange buffer): it can be
    # captured in a CUDA graph and replayed; every replayed step must reproduce the single-GPU loss
    from deep_staple_b200 import _binding as b

    lib = b.load()
    ex = dd.FusedExchange.get(B, C, dev)
    V = shape[0] * shape[1] * shape[2]
    z_l, t_l = logits[sl].contiguous(), target[sl].contiguous()
    bare_g = emb_w.reshape(-1).to(dev)[idx].contiguous()
    fixed_g = fixed[idx].contiguous()
    scale = torch.tensor(128.0, device=dev)
    dz = torch.empty_like(z_l)
    stats = torch.zeros((B, C, 4), dtype=torch.float64, device=dev)
    loss_out = torch.zeros(2, device=dev)
    d_bare = torch.zeros(B, device=dev)
    per_inst = torch.zeros((B, 4), device=dev)
    ws = torch.zeros((int(lib.ds_dp_loss_workspace_bytes(B_local, C)) + 7) // 8, dtype=torch.int64, device=dev)
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()

    def launch(stream):
        rc = lib.ds_dp_loss_fwd_bwd_exchange(
            z_l.data_ptr(), b.DS_F32, t_l.data_ptr(), b.DS_I64, None, B_local, C, V, bare_g.data_ptr(),
            fixed_g.data_ptr(), B, rank * B_local, cw.data_ptr(), 1, scale.data_ptr(), dz.data_ptr(), stats.data_ptr(),
            loss_out.data_ptr(), d_bare.data_ptr(), per_inst.data_ptr(), ex.peer_bases.data_ptr(), ex.rank, ex.world,
            0, ws.data_ptr(), ws.numel() * 8, stream)
        assert rc == 0, lib.ds_last_error_string()

    torch.cuda.synchronize()
    dist.barrier()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for _ in range(3):
                launch(side.cuda_stream)
    for rep in range(3):
        loss_out.zero_()
        graph.replay()
        torch.cuda.synchronize()
        close(loss_out[0], ref.dp_loss, 1e-6, f"graph replay {rep} dp_loss")
        close(loss_out[1], ref.ce_loss, 1e-6, f"graph replay {rep} ce_loss")
        assert torch.equal(stats, ref.inst_stats), f"graph replay {rep}: statistics"
        close(dz, z1.grad[sl], 1e-6, f"graph replay {rep} logits gradient")
    dist.barrier()

    # the global label histogram (normaliser of the class-weighted CE under data parallelism)
    counts = dd.global_class_counts(out.inst_stats[sl, :, 1].sum(0))
    assert torch.equal(counts, torch.bincount(target.reshape(-1), minlength=C).double())

    # ---- target-sharded consensus -------------------------------------------------------------------
    T, N, S = 5, 10, (16, 16, 32)
    gen = torch.Generator().manual_seed(1)
    labels = (torch.rand((T, N, *S), generator=gen) < 0.4).to(torch.uint8).to(dev)
    dps = torch.randn((T, N), generator=gen).to(dev)
    lo, hi = dd.shard_range(T, rank, world)
    ref_c, ref_t = ds.consensus_batched(labels, dps, count_near_ties=True)
    for fused in (False, True):  # NCCL all-gather after the vote / stores into every rank's buffer from inside the vote
        full, ties = dd.sharded_consensus(labels[lo:hi], dps[lo:hi], T, gather=True, count_near_ties=True, fused=fused)
        assert torch.equal(full, ref_c) and int(ties) == int(ref_t), f"sharded consensus, fused={fused}"
    # the fused gather again: int64 label stacks (uint8 on the wire, widened on arrival), the staged uint8 kernel forced,
    # ragged shards (T % world != 0 whenever world does not divide 5), repeated calls into the same symmetric buffers
    lab64 = labels.long()
    for rep in range(3):
        full64, _ = dd.sharded_consensus(lab64[lo:hi], dps[lo:hi] + rep, T, fused=True)
        ref64 = ds.consensus_batched(lab64, dps + rep)
        assert full64.dtype == torch.int64 and torch.equal(full64, ref64), f"fused gather int64 rep {rep}"
        full8, _ = dd.sharded_consensus(lab64[lo:hi], dps[lo:hi] + rep, T, fused=True, out_dtype=torch.uint8)
        assert full8.dtype == torch.uint8 and torch.equal(full8.long(), ref64)
    os.environ["DS_VOTE_STG_MIN_TILES"] = "0"
    full, _ = dd.sharded_consensus(labels[lo:hi], dps[lo:hi], T, fused=True)
    assert torch.equal(full, ref_c), "fused gather through the staged uint8 kernel"
    del os.environ["DS_VOTE_STG_MIN_TILES"]
    T2 = world + 1 if world > 1 else 3  # more ranks than full shares: some ranks own one target, rank 0 two
    lab2 = (torch.rand((T2, 3, 8, 8, 16), generator=gen) < 0.5).to(torch.uint8).to(dev)
    dps2 = torch.randn((T2, 3), generator=gen).to(dev)
    lo2, hi2 = dd.shard_range(T2, rank, world)
    full2, _ = dd.sharded_consensus(lab2[lo2:hi2], dps2[lo2:hi2], T2, fused=True)
    assert torch.equal(full2, ds.consensus_batched(lab2, dps2))
    lo3, hi3 = dd.shard_range(1, rank, world)  # one target: every rank but the first owns nothing
    full3, _ = dd.sharded_consensus(lab2[:1][lo3:hi3], dps2[:1][lo3:hi3], 1, fused=True)
    assert torch.equal(full3, ds.consensus_batched(lab2[:1], dps2[:1]))
    if rank == 0:
        print("FUSED_GATHER_OK", world)
    dist.barrier()
    if rank == 0:
        print("DIST_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
