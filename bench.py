#!/usr/bin/env python
"""bench.py -- the hot path of BASELINE.json on B200, one JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload loss|vote]

Workload (``config.workload``), BASELINE.json configs[1]: the data-parameter-weighted loss, fused
forward + backward with risk regularisation, fixed weighting, class-weighted CE value and hard-Dice
statistics, batch 8 x 2 classes x 128^3 per GPU, fp32 logits, int64 labels (the reference's dtypes).
A step = one pass of the hot path over one batch of synthetic input.

* ``value``      Gvoxel-candidates/s, inputs resident in HBM, the C-ABI call
                 ``ds_dp_loss_fwd_bwd`` per step (at N > 1 plus the NCCL all-gather of the
                 per-instance statistics and the replicated epilogue ``ds_dp_loss_finalize``).
                 Inputs rotate over 3 buffer sets (1.2 GB) so nothing is served from the 126 MB L2.
* ``roofline``   algorithmic bytes of the fused kernel (24 B per voxel-instance: 2 fp32 logits + one
                 int64 label read, 2 fp32 gradients written) / its average launch duration measured
                 with CUDA events on the launching stream, against MEASURED_PEAKS.json.
* ``e2e``        the same metric through the public Python API (``dp_weighted_loss`` + backward) with
                 HOST pinned inputs: per step H2D of logits + labels, D2H of the loss and the
                 data-parameter gradient, copies double-buffered against the kernel.
* ``cpu_baseline``  the oracle (restatement of main_deep_staple.py:716,738-762 on torch CPU) timed on
                 this box's host cores on a bounded sample (rank 0, N = 1 only).
* ``--impl reference``  times that same CPU path as the reference arm.
* ``vote``       (extra key) the consensus weighted vote on a cfg3 shard, with its own roofline.

``--workload vote`` makes the consensus vote (configs[2], sharded by target) the timed workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Gvoxel-candidates/s"
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the fused loss kernel on this workload,
# from the `ncu --set full` capture summarised in profiles/r01_loss_fwdbwd_ncu_summary.txt
# (268.5 MB read + 95.95 MB written; the rest of the 134 MB of gradients is still dirty in L2 at kernel end)
NCU_TRAFFIC_BYTES_PER_LAUNCH = 364_468_736

LOSS_CFG = dict(B=8, C=2, S=(128, 128, 128), M=600, grad_scale=65536.0)
VOTE_CFG = dict(T_global=100, N=30, S=(128, 128, 128))


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks during the timed regions
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover - depends on the box
            log(f"[clocks] NVML unavailable: {e}")

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self._h).gpu
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.samples.append((float(mhz), int(util)))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self._h is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        busy = [m for m, u in self.samples if u > 0] or [m for m, _ in self.samples]
        return {
            "sm_mhz": statistics.median(busy) if busy else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# ---------------------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------------------
def make_loss_inputs(device, seed, cfg=LOSS_CFG, B=None):
    B = B or cfg["B"]
    C, S, M = cfg["C"], cfg["S"], cfg["M"]
    gen = torch.Generator(device=device).manual_seed(seed)
    logits = 2 * torch.randn((B, C, *S), generator=gen, device=device)
    target = (torch.rand((B, *S), generator=gen, device=device) < 0.05).long()
    emb = torch.randn((M,), generator=gen, device=device)
    idx = torch.randperm(M, generator=gen, device=device)[:B]
    gt_num = torch.randint(0, 50000, (M,), generator=gen, device=device).float()
    fixed = (gt_num + math.e).log() + math.e
    counts = torch.bincount(target.reshape(-1), minlength=C).float()
    cw = 1 / counts.pow(0.35)
    cw = cw / cw.mean()
    return dict(logits=logits, target=target, emb=emb, idx=idx, fixed=fixed, cw=cw)


def ellipsoid_labels(T, N, S, device, seed, dtype):
    """cfg1(ii)/cfg3 labels: one ellipsoid per candidate, centre jittered around the volume centre."""
    gen = torch.Generator(device=device).manual_seed(seed)
    zz, yy, xx = torch.meshgrid(*[torch.arange(s, device=device, dtype=torch.float32) for s in S], indexing="ij")
    out = torch.empty((T, N, *S), dtype=dtype, device=device)
    centre = torch.tensor(S, device=device, dtype=torch.float32) / 2
    for t in range(T):
        c = centre + 3.0 * torch.randn((N, 3), generator=gen, device=device)
        r = 6.0 + 8.0 * torch.rand((N, 3), generator=gen, device=device)
        for n in range(N):
            out[t, n] = ((((zz - c[n, 0]) / r[n, 0]) ** 2 + ((yy - c[n, 1]) / r[n, 1]) ** 2 + ((xx - c[n, 2]) / r[n, 2]) ** 2) <= 1.0).to(dtype)
    return out


# ---------------------------------------------------------------------------------------------
# the reference arm / cpu baseline: the oracle on host cores
# ---------------------------------------------------------------------------------------------
def cpu_loss_baseline(steps, warmup, budget_s=120.0):
    """Times oracle.dp_loss.run_loss (restatement of main_deep_staple.py:716,738-762, torch CPU, all
    host threads) on a bounded sample of cfg2.  Returns (Gvox/s, ms/step, cores, sample string)."""
    from oracle import dp_loss as ol

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = LOSS_CFG
    full = make_loss_inputs("cpu", 0)

    def run(B, D):
        z = full["logits"][:B, :, :D].contiguous()
        t = full["target"][:B, :D].contiguous()
        t0 = time.perf_counter()
        ol.run_loss(z, t, full["emb"].reshape(-1, 1), full["idx"][:B], full["fixed"], full["cw"], True,
                    cfg["grad_scale"], True)
        return time.perf_counter() - t0

    t_probe = run(1, 32) * 4  # one instance, from a quarter-depth probe
    n = max(1, steps + warmup)
    per_step_budget = budget_s / n
    B = int(max(1, min(cfg["B"], per_step_budget // max(t_probe, 1e-6))))
    D = cfg["S"][0]
    if B == 1 and t_probe > per_step_budget:
        D = int(max(8, min(cfg["S"][0], cfg["S"][0] * per_step_budget / t_probe)))
    for _ in range(warmup):
        run(B, D)
    times = [run(B, D) for _ in range(max(1, steps))]
    vox = B * D * cfg["S"][1] * cfg["S"][2]
    sec = sum(times) / len(times)
    sample = (f"{B} of {cfg['B']} instances x {cfg['C']} classes x {D}x{cfg['S'][1]}x{cfg['S'][2]} voxels per step, "
              f"fwd+bwd incl. dlogits, {len(times)} timed steps")
    return vox / sec / 1e9, sec * 1e3, cores, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, cores, sample = cpu_loss_baseline(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gvoxel-candidates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": loss_config_dict(1),
        "cpu_baseline": {"value": value, "unit": "Gvoxel-candidates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gvoxel-candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = torch CPU path of main_deep_staple.py:716,738-762 (oracle port; the reference code is "
                "inline in train_DL and cannot be imported), all host threads",
    }
    emit(line)


def loss_config_dict(n_gpus, exchange="inside the loss kernel over NVLink peer memory", launch="n/a"):
    c = LOSS_CFG
    return {
        "workload": "dp-weighted CE loss fwd+bwd (risk regularisation, fixed weighting, class-weighted CE value, "
                    "hard-Dice statistics), batch 8 x 2 classes x 128^3 per GPU (BASELINE.json configs[1])",
        "batch_per_gpu": c["B"], "global_batch": c["B"] * n_gpus, "classes": c["C"], "volume": list(c["S"]),
        "logits_dtype": "float32", "label_dtype": "int64", "grad_scale": c["grad_scale"],
        "parallelism": f"dp{n_gpus} (instances sharded; per-instance statistics exchanged {exchange})" if n_gpus > 1 else "single GPU",
        "l2": "inputs rotate over 3 buffer sets of 403 MB (> 126 MB L2); no reuse between steps",
        "launch": launch,
    }


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
class LossStep:
    """Preallocated buffers + the raw C-ABI call of one step (what `value` and `roofline` time)."""

    def __init__(self, device, n_sets, rank, world, exchange="fused"):
        import deep_staple_b200 as ds
        from deep_staple_b200 import _binding as b

        self.ds, self.b, self.lib = ds, b, b.load()
        self.device, self.rank, self.world = device, rank, world
        c = LOSS_CFG
        self.B, self.C = c["B"], c["C"]
        self.V = c["S"][0] * c["S"][1] * c["S"][2]
        self.B_total = self.B * world
        self.sets = [make_loss_inputs(device, 1000 * rank + i) for i in range(n_sets)]
        self.dlogits = [torch.empty_like(s["logits"]) for s in self.sets]
        s0 = self.sets[0]
        # data parameters / fixed weights of the GLOBAL batch (identical on every rank: same seed)
        gen = torch.Generator(device=device).manual_seed(4242)
        self.bare = torch.randn(self.B_total, generator=gen, device=device)
        self.fixed = (torch.randint(0, 50000, (self.B_total,), generator=gen, device=device).float() + math.e).log() + math.e
        self.cw = s0["cw"].contiguous()
        self.scale = torch.tensor(c["grad_scale"], device=device)
        self.stats = torch.zeros((self.B_total, self.C, 4), dtype=torch.float64, device=device)
        self.loss_out = torch.zeros(2, device=device)
        self.d_bare = torch.zeros(self.B_total, device=device)
        self.per_inst = torch.zeros((self.B_total, 4), device=device)
        nbytes = int(self.lib.ds_dp_loss_workspace_bytes(self.B, self.C))
        self.ws = torch.zeros((nbytes + 7) // 8, dtype=torch.int64, device=device)
        self.stream = torch.cuda.current_stream(device).cuda_stream
        self.i = 0
        self.exchange = None
        if world > 1 and exchange == "fused":
            from deep_staple_b200.distributed import FusedExchange

            self.exchange = FusedExchange.get(self.B_total, self.C, device)

    def kernel(self):
        b, k = self.b, self.i % len(self.sets)
        self.i += 1
        s = self.sets[k]
        rc = self.lib.ds_dp_loss_fwd_bwd(
            s["logits"].data_ptr(), b.DS_F32, s["target"].data_ptr(), b.DS_I64, None, self.B, self.C, self.V,
            self.bare.data_ptr(), self.fixed.data_ptr(), self.B_total, self.rank * self.B, self.cw.data_ptr(), 1,
            self.scale.data_ptr(), self.dlogits[k].data_ptr(), self.stats.data_ptr(), self.loss_out.data_ptr(),
            self.d_bare.data_ptr(), self.per_inst.data_ptr(), self.ws.data_ptr(), self.ws.numel() * 8, self.stream)
        if rc:
            b.check(rc, "ds_dp_loss_fwd_bwd")

    def fused_step(self):
        """data-parallel step as ONE launch: statistics exchange over NVLink peer memory + epilogue in the kernel tail"""
        b, k, ex = self.b, self.i % len(self.sets), self.exchange
        self.i += 1
        s = self.sets[k]
        rc = self.lib.ds_dp_loss_fwd_bwd_exchange(
            s["logits"].data_ptr(), b.DS_F32, s["target"].data_ptr(), b.DS_I64, None, self.B, self.C, self.V,
            self.bare.data_ptr(), self.fixed.data_ptr(), self.B_total, self.rank * self.B, self.cw.data_ptr(), 1,
            self.scale.data_ptr(), self.dlogits[k].data_ptr(), self.stats.data_ptr(), self.loss_out.data_ptr(),
            self.d_bare.data_ptr(), self.per_inst.data_ptr(), ex.peer_bases.data_ptr(), ex.rank, ex.world,
            ex.next_epoch(), self.ws.data_ptr(), self.ws.numel() * 8, self.stream)
        if rc:
            b.check(rc, "ds_dp_loss_fwd_bwd_exchange")

    def step(self):
        if self.exchange is not None:
            self.fused_step()
            return
        self.kernel()
        if self.world > 1:
            import torch.distributed as dist

            mine = self.stats[self.rank * self.B:(self.rank + 1) * self.B]
            dist.all_gather_into_tensor(self.stats, mine)
            rc = self.lib.ds_dp_loss_finalize(
                self.stats.data_ptr(), self.B_total, self.C, self.V, self.bare.data_ptr(), self.fixed.data_ptr(),
                self.cw.data_ptr(), 1, self.scale.data_ptr(), self.loss_out.data_ptr(), self.d_bare.data_ptr(),
                self.per_inst.data_ptr(), self.stream)
            if rc:
                self.b.check(rc, "ds_dp_loss_finalize")


def timed(fn, steps, warmup, barrier=None):
    """W warm-up steps, then K steps between CUDA events on the current stream; returns seconds."""
    for _ in range(warmup):
        fn()
    if barrier:
        barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    return e0.elapsed_time(e1) / 1e3


def graph_chunk(steps, cap=256):
    """Steps per captured graph: `steps` itself when small, else its largest divisor <= cap (0: none useful)."""
    if steps <= cap:
        return steps
    best = max(d for d in range(1, cap + 1) if steps % d == 0)
    return best if best >= 16 else 0


def timed_graph(step_obj, fn, steps, warmup, barrier=None):
    """Like timed(), but the K steps are replayed from a CUDA graph (steps // chunk replays of a chunk-step
    graph captured on a side stream): the launches are enqueued by the GPU front end instead of one by one
    from Python, which removes ~2 us of inter-kernel gap per step.  Still exactly K steps between two CUDA
    events on the replaying stream, after W warm-up steps.  Returns (seconds, launches_per_step)."""
    import deep_staple_b200 as ds

    chunk = graph_chunk(steps)
    if chunk == 0:
        return timed(fn, steps, warmup, barrier), None
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    main_stream = step_obj.stream
    step_obj.stream = side.cuda_stream
    try:
        with torch.cuda.stream(side):
            n0 = ds.launch_count()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(chunk):
                    fn()
            per_step = (ds.launch_count() - n0) / chunk
    finally:
        step_obj.stream = main_stream
    graph.replay()  # one untimed replay: graph upload
    if barrier:
        barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps // chunk):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    return e0.elapsed_time(e1) / 1e3, per_step


def e2e_loss(device, steps, warmup):
    """Public API with HOST inputs: H2D (pinned) of logits + labels, fused kernel via
    dp_weighted_loss + backward, D2H of loss and data-parameter gradient; copies for step i+1 are
    issued on a second stream while step i computes."""
    import deep_staple_b200 as ds

    c = LOSS_CFG
    host = make_loss_inputs("cpu", 7)
    h_logits = [host["logits"].pin_memory(), host["logits"].clone().pin_memory()]
    h_target = [host["target"].pin_memory(), host["target"].clone().pin_memory()]
    d_logits = [torch.empty_like(host["logits"], device=device) for _ in range(2)]
    d_target = [torch.empty_like(host["target"], device=device) for _ in range(2)]
    bare0 = host["emb"][host["idx"]].to(device)
    fixed = host["fixed"][host["idx"]].to(device)
    cw = host["cw"].to(device)
    h_out = torch.empty(1 + c["B"], dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device)
    main = torch.cuda.current_stream(device)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    h2d = host["logits"].numel() * 4 + host["target"].numel() * 8
    d2h = h_out.numel() * 4

    def upload(k):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[k])
            d_logits[k].copy_(h_logits[k], non_blocking=True)
            d_target[k].copy_(h_target[k], non_blocking=True)
            ready[k].record(copy_stream)

    def compute(k):
        main.wait_event(ready[k])
        z = d_logits[k].requires_grad_(True)
        bare = bare0.clone().requires_grad_(True)
        loss = ds.dp_weighted_loss(z, d_target[k], bare, fixed, True, grad_scaler=c["grad_scale"], class_weights=cw)
        (loss * c["grad_scale"]).backward()
        h_out[:1].copy_(loss.detach().reshape(1), non_blocking=True)
        h_out[1:].copy_(bare.grad, non_blocking=True)
        z.grad = None
        d_logits[k].requires_grad_(False)
        freed[k].record(main)

    for k in range(2):
        freed[k].record(main)
    n = warmup + steps
    upload(0)
    t_start = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(n):
        if i == warmup:
            torch.cuda.synchronize()
            upload(i % 2)  # the timed region starts with its own first upload
            e0.record(main)
            t_start = time.perf_counter()
        if i + 1 < n and i + 1 != warmup:
            upload((i + 1) % 2)
        compute(i % 2)
    e1.record(main)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_start
    sec = max(e0.elapsed_time(e1) / 1e3, wall)
    return sec, h2d, d2h, float(h_out[0])


def vote_section(device, rank, world, steps, warmup, peak):
    """Consensus vote on this rank's cfg3 shard (targets sharded contiguously; no collective in the
    timed region).  uint8 labels are the compact opt-in dtype, int64 the reference's."""
    import deep_staple_b200 as ds
    from deep_staple_b200.distributed import shard_range

    c = VOTE_CFG
    lo, hi = shard_range(c["T_global"], rank, world)
    T = min(hi - lo, 13)  # 13 targets = the 8-GPU shard; 30 x 128^3 int64 x 13 = 6.5 GB, far beyond L2
    V = c["S"][0] * c["S"][1] * c["S"][2]
    out = {}
    gen = torch.Generator(device=device).manual_seed(4)
    dps = 2 * torch.randn((T, c["N"]), generator=gen, device=device)
    for name, dt, sz in (("int64", torch.int64, 8), ("uint8", torch.uint8, 1)):
        labels = ellipsoid_labels(T, c["N"], c["S"], device, 1000 + lo, dt)
        res = torch.empty((T, *c["S"]), dtype=dt, device=device)
        fn = lambda: ds.consensus_batched(labels, dps, out=res)
        sec = timed(fn, steps, warmup)
        bytes_per_launch = T * V * (c["N"] * sz + sz)
        out[name] = {
            "gvoxel_candidates_per_s": T * c["N"] * V * steps / sec / 1e9,
            "ms_per_launch": sec / steps * 1e3,
            "targets": T, "candidates": c["N"],
            "roofline": {"bound": "hbm", "achieved": bytes_per_launch * steps / sec / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bytes_per_launch * steps / sec / 1e9 / peak},
        }
        del labels, res
    return out


_REAL_STDOUT = None


def quiet_stdout():
    """Everything but the final JSON line goes to stderr: libraries (NCCL's version banner, c10d) write to
    file descriptor 1 directly, and the contract is ONE JSON line on stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--skip-vote", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--launch", default="graph", choices=["graph", "stream"],
                    help="timed steps replayed from a CUDA graph (default) or launched one by one from Python")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N > 1: statistics exchange inside the loss kernel over NVLink peer memory, or NCCL all-gather")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible and deep_staple_b200 has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)
    barrier = (lambda: dist.barrier()) if dist else None

    import deep_staple_b200 as ds

    peak, peak_src = peak_hbm()
    step = LossStep(device, 3, rank, world, args.exchange)
    c = LOSS_CFG
    vox_per_step = step.B * step.V  # voxel-instances per GPU per step

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step.step()
    use_graph = args.launch == "graph" and (world == 1 or args.exchange == "fused")
    launches0 = ds.launch_count()
    if use_graph:
        sec, per_step = timed_graph(step, step.step, args.steps, 0, barrier)
        launches = round(per_step * args.steps) if per_step is not None else ds.launch_count() - launches0
        use_graph = per_step is not None
    else:
        sec = timed(step.step, args.steps, 0, barrier)
        launches = ds.launch_count() - launches0
    if dist:
        dist.barrier()  # nobody leaves the fused-exchange phase while a peer still waits on its flags
    if dist:
        tmax = torch.tensor([sec], device=device, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        sec = float(tmax.item())
    # the dominant kernel alone (no collective, no epilogue launch): average launch duration
    if use_graph:
        ksec, _ = timed_graph(step, step.kernel, args.steps, max(3, args.warmup // 2))
    else:
        ksec = timed(step.kernel, args.steps, max(3, args.warmup // 2))
    e2e = None
    if not args.skip_e2e:
        del step.sets[1:], step.dlogits[1:]
        torch.cuda.empty_cache()
        esec, h2d, d2h, _ = e2e_loss(device, args.e2e_steps, 3)
        if dist:
            tmax = torch.tensor([esec], device=device, dtype=torch.float64)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            esec = float(tmax.item())
        e2e = {"value": world * vox_per_step * args.e2e_steps / esec / 1e9, "unit": "Gvoxel-candidates/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
               "api": "deep_staple_b200.dp_weighted_loss(...) + backward, pinned host inputs, double-buffered copies"}
    clocks = sampler.stop()
    del step
    torch.cuda.empty_cache()

    vote = None
    if not args.skip_vote:
        vote = vote_section(device, rank, world, max(3, min(args.steps, 20)), 3, peak)
        if dist:
            agg = torch.tensor([vote["int64"]["gvoxel_candidates_per_s"], vote["uint8"]["gvoxel_candidates_per_s"]],
                               device=device, dtype=torch.float64)
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
            vote["all_ranks_gvoxel_candidates_per_s"] = {"int64": float(agg[0]), "uint8": float(agg[1])}

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        v, ms, cores, sample = cpu_loss_baseline(3, 1, budget_s=25.0)
        cpu = {"value": v, "unit": "Gvoxel-candidates/s", "cores": cores, "kind": "port", "sample": sample,
               "ms_per_step": ms, "torch": torch.__version__}

    if rank == 0:
        alg_bytes = vox_per_step * (2 * c["C"] * 4 + 8)  # 24 B per voxel-instance
        achieved = alg_bytes * args.steps / ksec / 1e9
        line = {
            "metric": METRIC, "value": world * vox_per_step * args.steps / sec / 1e9, "unit": "Gvoxel-candidates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": loss_config_dict(world, "inside the loss kernel over NVLink peer memory" if args.exchange == "fused"
                                       else "by NCCL all-gather + epilogue launch",
                                       f"CUDA graph of {graph_chunk(args.steps)} steps, replayed" if use_graph
                                       else "one C-ABI call per step from the host"),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH, "traffic_source": "profiles/r01_loss_fwdbwd_ncu_summary.txt",
                         "kernel": "ds::loss_kernel<float,int64,2,4,true,true>",
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_us": ksec / args.steps * 1e6,
                         "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "vote": vote,
        }
        emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
