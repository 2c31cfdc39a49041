import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, deep_staple_b200 as ds
dev = torch.device("cuda:0")
S = (128, 128, 128); V = 128 ** 3
for T, N in ((1, 10), (1, 30), (4, 10)):
    for name, dt, sz in (("i64", torch.int64, 8), ("u8", torch.uint8, 1)):
        nsets = max(2, int(1.5e9 // (T * N * V * sz)))
        g = torch.Generator(device=dev).manual_seed(1)
        sets = [((torch.rand((T, N, *S), generator=g, device=dev) < 0.3).to(dt), torch.empty((T, *S), dtype=dt, device=dev)) for _ in range(min(nsets, 16))]
        dps = 2 * torch.randn((T, N), generator=g, device=dev)
        st = {"i": 0}
        def fn():
            lab, out = sets[st["i"] % len(sets)]; st["i"] += 1
            ds.consensus_batched(lab, dps, out=out)
        alg = T * V * (N + 1) * sz
        for per_sm, waves in ((4, 1), (4, 2), (4, 4), (4, 8), (4, 16), (8, 1), (8, 2)):
            os.environ["DS_VOTE_CTAS_PER_SM"], os.environ["DS_VOTE_WAVES"] = str(per_sm), str(waves)
            sec = bench.timed(fn, 200, 5) / 200
            print(f"T={T} N={N} {name} per_sm={per_sm} waves={waves}: {sec*1e6:7.1f} us {alg/sec/1e9:7.0f} GB/s", flush=True)
        del sets
