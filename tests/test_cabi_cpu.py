"""CPU-only checks of the C-ABI boundary: the shared library loads, exports every symbol that
include/deep_staple_b200.h declares, validates its arguments, and refuses to compute without a
B200 (no CPU fallback anywhere)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "deep_staple_b200.h")


def declared_functions():
    with open(HEADER) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|size_t|uint64_t|const char\*)\s+(ds_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = n
    return decls


@pytest.fixture(scope="module")
def lib():
    import deep_staple_b200 as ds

    return ds.load()


def test_header_declares_the_expected_entry_points():
    decls = declared_functions()
    for name in ("ds_consensus_vote", "ds_consensus_vote_sparse", "ds_consensus_vote_multiclass", "ds_dp_loss_fwd_bwd",
                 "ds_dp_loss_finalize", "ds_weighted_ce_fwd_bwd", "ds_dp_loss_backward_logits", "ds_rescale_inplace",
                 "ds_label_stats", "ds_dp_loss_workspace_bytes", "ds_version", "ds_last_error_string",
                 "ds_device_check", "ds_launch_count"):
        assert name in decls, name


def test_library_exports_every_declared_symbol_and_binding_matches(lib):
    from deep_staple_b200 import _binding as b

    decls = declared_functions()
    assert set(decls) == set(b.SIGNATURES), (set(decls) ^ set(b.SIGNATURES))
    for name, nargs in decls.items():
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
        assert len(b.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args, binding {len(b.SIGNATURES[name][1])}"


def test_version_and_workspace_size(lib):
    assert lib.ds_version() == 100
    assert lib.ds_dp_loss_workspace_bytes(0, 2) == 0
    small, big = lib.ds_dp_loss_workspace_bytes(8, 2), lib.ds_dp_loss_workspace_bytes(64, 8)
    assert 0 < small < big
    assert small % 8 == 0


def test_argument_validation_happens_before_any_device_work(lib):
    fake = ctypes.c_void_p(0x1000)  # never dereferenced: all of these fail validation first
    assert lib.ds_consensus_vote(None, 3, fake, 1, 1, 2, 8, fake, 3, None, None) == -1  # DS_ERR_NULL
    assert b"required" in lib.ds_last_error_string()
    assert lib.ds_consensus_vote(fake, 0, fake, 1, 1, 2, 8, fake, 3, None, None) == -2  # DS_ERR_DTYPE (f32 labels)
    assert lib.ds_consensus_vote(fake, 3, fake, 1, 0, 2, 8, fake, 3, None, None) == -3  # DS_ERR_SHAPE (T = 0)
    assert lib.ds_consensus_vote(fake, 3, fake, 1, 1, 70000, 8, fake, 3, None, None) == -3  # N > DS_MAX_CANDIDATES
    assert lib.ds_dp_loss_fwd_bwd(fake, 0, fake, 3, None, 2, 9, 64, fake, None, 2, 0, None, 1, None, None, fake, fake,
                                  fake, fake, fake, 1 << 20, None) == -3  # C = 9
    assert lib.ds_dp_loss_fwd_bwd(fake, 7, fake, 3, None, 2, 2, 64, fake, None, 2, 0, None, 1, None, None, fake, fake,
                                  fake, fake, fake, 1 << 20, None) == -2  # unknown logits dtype
    assert lib.ds_rescale_inplace(fake, 0, 0, fake, fake, None) == -3
    # fused vote + gather: the peer table, the world size and the row range are checked before anything else
    table = (ctypes.c_void_p * 2)(0x1000, 0x2000)
    assert lib.ds_consensus_vote_gather(fake, 4, fake, 1, 1, 2, 64, 0, 2, None, 2, 4, None, None) == -1
    assert lib.ds_consensus_vote_gather(fake, 4, fake, 1, 1, 2, 64, 0, 2, table, 0, 4, None, None) == -6   # world_size 0
    assert lib.ds_consensus_vote_gather(fake, 4, fake, 1, 1, 2, 64, 0, 2, table, 17, 4, None, None) == -6  # > DS_MAX_PEERS
    assert lib.ds_consensus_vote_gather(fake, 4, fake, 1, 2, 2, 64, 1, 2, table, 2, 4, None, None) == -6   # rows 1..2 of 2
    assert lib.ds_consensus_vote_gather(fake, 4, fake, 1, 0, 2, 64, 2, 2, table, 2, 4, None, None) == 0    # an empty shard
    assert lib.ds_consensus_vote_gather(fake, 0, fake, 1, 1, 2, 64, 0, 2, table, 2, 4, None, None) == -2   # f32 labels
    hole = (ctypes.c_void_p * 2)(0x1000, None)
    if torch.cuda.is_available():
        assert lib.ds_consensus_vote_gather(fake, 4, fake, 1, 1, 2, 64, 0, 2, hole, 2, 4, None, None) == -1
    # host label packer
    assert lib.ds_host_narrow_labels(None, fake, 8, 1) == -1
    assert lib.ds_host_narrow_labels(fake, fake, -1, 1) == -3
    assert lib.ds_host_narrow_labels(fake, fake, 0, 1) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_a_gpu(lib):
    import deep_staple_b200 as ds

    fake = ctypes.c_void_p(0x1000)
    assert lib.ds_device_check() == -5  # DS_ERR_DEVICE
    # valid arguments, but no B200: the entry points refuse instead of computing something on the host
    assert lib.ds_consensus_vote(fake, 3, fake, 1, 1, 2, 8, fake, 3, None, None) == -5
    assert lib.ds_label_stats(fake, 3, None, 3, 1, 2, 8, fake, fake, 1 << 20, None) == -5
    table = (ctypes.c_void_p * 2)(0x1000, 0x2000)
    assert lib.ds_consensus_vote_gather(fake, 4, fake, 1, 1, 2, 64, 0, 2, table, 2, 4, None, None) == -5
    assert lib.ds_launch_count() == 0
    lbl = torch.zeros((4, 4, 4), dtype=torch.long)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ds.calc_dp_consensus([lbl, lbl], [0.1, 0.2])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ds.dp_weighted_loss(torch.zeros((2, 2, 4, 4)), torch.zeros((2, 4, 4), dtype=torch.long), torch.zeros(2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ds.consensus_batched(torch.zeros((1, 2, 4), dtype=torch.long), torch.zeros((1, 2)))


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure; nothing under deep_staple_b200/ may import it."""
    pkg = os.path.join(ROOT, "deep_staple_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_snapshot_format_round_trip(tmp_path):
    """main_deep_staple.py:963-1037 / consensus.ipynb:330-377: keys, dtypes, ascending order, sparse stacks."""
    import deep_staple_b200 as ds

    torch.manual_seed(0)
    shape = (4, 6, 8)
    ids = ["126l:m102l", "126l:m103r", "127r:m102l"]
    dps = torch.tensor([0.7, -1.2, 0.1])
    mod = [(torch.rand(shape) < 0.3).long() for _ in ids]
    snap = ds.build_snapshot(dps, ids, [5, 9, 2], modified_labels=mod, labels=mod,
                             train_predictions=[m.unsqueeze(0) for m in mod], disturb_flags=[False, True, False],
                             image_paths=["a", "b", "c"], label_paths=["x", "y", "z"])
    path = ds.save_snapshot(snap, tmp_path / "out" / "train_label_snapshot.pth")
    back = ds.load_snapshot(path)
    assert list(back["d_ids"]) == ["126l:m103r", "127r:m102l", "126l:m102l"]  # sorted ascending by data parameter
    assert back["data_parameters"].shape == (3, 1) and back["data_parameters"].dtype == torch.float32
    assert back["data_parameters"].requires_grad, "requires_grad must survive save/load like in the reference (SURVEY a10)"
    assert torch.all(back["data_parameters"][1:] >= back["data_parameters"][:-1])
    assert back["dataset_idxs"].tolist() == [9, 2, 5] and back["dataset_idxs"].dtype == torch.int64
    assert back["disturb_flags"] == (True, False, False)
    assert back["modified_labels"].is_sparse and back["modified_labels"].shape == (3, *shape)
    assert back["train_predictions"].shape == (3, 1, *shape)
    assert back["modified_labels"][0].is_sparse and back["data_parameters"][0].shape == (1,)
    assert torch.equal(back["modified_labels"][0].to_dense(), mod[1])
    groups = ds.group_snapshot(back)
    assert list(groups.keys()) == ["126l", "127r"] and set(groups["126l"].keys()) == {"expert_label", "prediction", "103r", "102l"}


def test_bench_graph_chunking_times_exactly_k_steps():
    """bench.py replays its K timed steps from a CUDA graph of `chunk` steps: chunk must divide K exactly
    (the contract is EXACTLY K steps between the events), small K is captured whole, and awkward K (no
    useful divisor) falls back to one launch per step."""
    import bench

    for k in (1, 2, 7, 50, 250, 256, 2000, 3000, 1024):
        c = bench.graph_chunk(k)
        assert c >= 1 and k % c == 0 and c <= max(k, 256) and (c == k or c <= 256), (k, c)
    assert bench.graph_chunk(2000) == 250
    assert bench.graph_chunk(257) == 0 and bench.graph_chunk(2 * 257) == 0  # primes beyond the cap: stream launches
