"""Host-side label packer (ds_host_narrow_labels, include/deep_staple_b200.h): runs on the CPU, no GPU needed.
Replaces the int64 label upload of main_deep_staple.py:687 -- must be lossless and must refuse what it cannot carry."""
import os
import subprocess
import sys

import pytest
import torch

import deep_staple_b200 as ds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 15, 16, 17, 31, 32, 33, 255, 4096, 262144 + 5, 3 * 262144 + 1])
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_narrow_matches_torch_cast(n, threads):
    gen = torch.Generator().manual_seed(n + threads)
    x = torch.randint(0, 256, (n,), generator=gen, dtype=torch.int64)
    out = ds.narrow_labels_host(x, threads=threads)
    assert out.dtype == torch.uint8 and out.shape == x.shape
    assert torch.equal(out, x.to(torch.uint8))


def test_narrow_unaligned_views_and_shapes():
    gen = torch.Generator().manual_seed(1)
    base = torch.randint(0, 4, (100_003,), generator=gen, dtype=torch.int64)
    staging = torch.empty(100_040, dtype=torch.uint8)
    for off in range(0, 19):
        src = base[off:]
        dst = staging[off % 7: off % 7 + src.numel()]
        assert torch.equal(ds.narrow_labels_host(src, out=dst), src.to(torch.uint8))
    vol = torch.randint(0, 3, (2, 5, 6, 7), generator=gen, dtype=torch.int64)
    assert torch.equal(ds.narrow_labels_host(vol.transpose(1, 2)), vol.transpose(1, 2).to(torch.uint8))  # non-contiguous input


@pytest.mark.parametrize("bad", [256, -1, 1 << 40, -(1 << 62)])
@pytest.mark.parametrize("pos", [0, 5, 31, 32, 99_998, 262144 + 3])
def test_narrow_refuses_labels_it_cannot_carry(bad, pos):
    x = torch.ones(262144 + 9, dtype=torch.int64)
    x[pos] = bad
    with pytest.raises(ValueError, match="outside"):
        ds.narrow_labels_host(x, threads=4)


def test_narrow_argument_errors():
    with pytest.raises(ValueError):
        ds.narrow_labels_host(torch.zeros(4, dtype=torch.int32))
    with pytest.raises(ValueError):
        ds.narrow_labels_host(torch.zeros(4, dtype=torch.int64), out=torch.zeros(5, dtype=torch.uint8))
    with pytest.raises(ValueError):
        ds.narrow_labels_host(torch.zeros(4, dtype=torch.int64), out=torch.zeros(4, dtype=torch.int8))


@pytest.mark.parametrize("isa", ["scalar", "avx2", "avx512"])
def test_every_isa_variant_the_cpu_supports(isa):
    """DS_HOST_ISA is read once per process: each variant runs in its own interpreter (falls back to the best
    supported variant when the CPU lacks the requested one -- the result must be right either way)."""
    code = (
        "import torch, deep_staple_b200 as ds\n"
        "g = torch.Generator().manual_seed(3)\n"
        "x = torch.randint(0, 256, (1_000_003,), generator=g, dtype=torch.int64)\n"
        "assert all(torch.equal(ds.narrow_labels_host(x[k:], threads=t), x[k:].to(torch.uint8)) for k in range(34) for t in (1, 4))\n"
        "x[777_777] = 300\n"
        "try:\n    ds.narrow_labels_host(x)\n    raise SystemExit(3)\nexcept ValueError:\n    pass\n"
    )
    env = dict(os.environ, DS_HOST_ISA=isa, PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]


def test_concurrent_callers_share_the_pool():
    import threading

    gen = torch.Generator().manual_seed(9)
    xs = [torch.randint(0, 200, (600_000 + 1000 * i,), generator=gen, dtype=torch.int64) for i in range(6)]
    outs = [None] * len(xs)

    def work(i):
        for _ in range(5):
            outs[i] = ds.narrow_labels_host(xs[i], threads=4)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(xs))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for x, o in zip(xs, outs):
        assert torch.equal(o, x.to(torch.uint8))


def test_forked_child_gets_its_own_pool():
    """A forked process (a DataLoader worker) inherits the pool object but none of its threads: it must start a pool
    of its own instead of waiting for helpers that do not exist.  Runs in a fresh interpreter: the parent uses the
    pool, forks, the child packs with helper threads and reports through its exit code."""
    code = (
        "import os, sys, torch, deep_staple_b200 as ds\n"
        "torch.set_num_threads(1)\n"  # torch's own OpenMP team does not survive a fork either: keep it out of the test
        "x = torch.arange(3_000_000, dtype=torch.int64) % 251\n"
        "ref = x.to(torch.uint8)\n"
        "assert torch.equal(ds.narrow_labels_host(x, threads=4), ref)\n"
        "pid = os.fork()\n"
        "if pid == 0:\n"
        "    ok = all(torch.equal(ds.narrow_labels_host(x, threads=4), ref) for _ in range(3))\n"
        "    os._exit(0 if ok else 7)\n"
        "_, status = os.waitpid(pid, 0)\n"
        "assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0, status\n"
        "assert torch.equal(ds.narrow_labels_host(x, threads=4), ref)\n"
    )
    res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True,
                         text=True, timeout=120)
    assert res.returncode == 0, res.stderr[-2000:]
