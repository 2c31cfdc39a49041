"""GPU parity: fused data-parameter loss / class-weighted CE (CUDA, through the C ABI) vs the
reference outputs in tests/golden/ and vs the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): loss within 1e-5 relative in fp32, gradients within 1e-4
relative (relative to the largest gradient magnitude of the tensor; fp16 gradients additionally carry
their own fp16 rounding, 2^-11 relative per element).
"""
import math

import numpy as np
import pytest
import torch
from torch import nn

from conftest import golden_names, load_golden
from oracle import dp_loss as ol

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4


def _close(a, b, rtol, what, atol=0.0):
    a = torch.as_tensor(a).double().cpu().reshape(-1)
    b = torch.as_tensor(b).double().cpu().reshape(-1)
    scale = float(b.abs().max()) if b.numel() else 0.0
    err = float((a - b).abs().max()) if b.numel() else 0.0
    assert err <= rtol * scale + atol, f"{what}: max abs err {err:.3e} vs scale {scale:.3e} (rtol {rtol})"
    return err / (scale + 1e-300)


def _close_per_instance(a, b, rtol, what, atol=0.0):
    """Every instance's gradient against ITS OWN magnitude: dlogits carries w_b / V, which differs by orders of
    magnitude between instances, so a tensor-wide max-norm check barely tests the small-weight instances."""
    a = torch.as_tensor(a).detach().double().cpu().flatten(1)
    b = torch.as_tensor(b).detach().double().cpu().flatten(1)
    err = (a - b).abs().amax(1)
    scale = b.abs().amax(1)
    bad = err > rtol * scale + atol
    assert not bool(bad.any()), (f"{what}: instances {bad.nonzero().flatten().tolist()} exceed rtol {rtol}: "
                                 f"err {err[bad].tolist()} vs own scale {scale[bad].tolist()}")


def _run_case(g, dev, label_dtype=torch.int64):
    import deep_staple_b200 as ds

    logits = torch.from_numpy(g["logits"]).to(dev)
    target = torch.from_numpy(g["target"].astype(np.int64)).to(label_dtype).to(dev)
    m = g["emb"].shape[0]
    emb = nn.Embedding(m, 1, sparse=True, _weight=torch.from_numpy(g["emb"]).clone().reshape(m, 1)).to(dev)
    idx = torch.from_numpy(g["idx"]).to(dev)
    fixed = torch.from_numpy(g["fixed"]).to(dev)[idx] if bool(g["use_fixed"]) else None
    cw = torch.from_numpy(g["cw"]).to(dev)
    scale = float(g["g"])
    z = logits.clone().requires_grad_(bool(g["need_dlogits"]))
    bare = emb(idx).squeeze()  # main_deep_staple.py:741
    out = ds.dp_weighted_loss(z, target, bare, fixed, bool(g["risk"]), grad_scaler=scale, class_weights=cw,
                              return_all=True)
    (out.dp_loss * scale).backward()  # scaler_dp.scale(dp_loss).backward(), :762
    return out, z, emb, idx, cw, logits, target, scale


@pytest.mark.parametrize("name", golden_names("loss"))
@pytest.mark.parametrize("label_dtype", [torch.int64, torch.uint8])
def test_golden_dp_loss(cuda_device, name, label_dtype):
    g = load_golden(name)
    out, z, emb, idx, cw, logits, target, scale = _run_case(g, cuda_device, label_dtype)
    half = g["logits"].dtype != np.float32
    _close(out.dp_loss, g["dp_loss"], LOSS_RTOL, "dp_loss")
    _close(out.ce_loss, g["ce_loss"], LOSS_RTOL, "ce_loss")
    grad = emb.weight.grad
    assert grad.is_sparse, "the data-parameter gradient must stay sparse COO (SparseAdam rejects dense grads)"
    rows = grad.coalesce().to_dense().reshape(-1)[idx]
    _close(rows, g["d_bare_weight"], GRAD_RTOL, "d_bare_weight", atol=1e-9 * scale)
    if bool(g["need_dlogits"]):
        assert z.grad.dtype == z.dtype and z.grad.shape == z.shape
        _close(z.grad, g["d_logits"], 2e-3 if half else GRAD_RTOL, "d_logits")
        # fp16 gradients: 2^-11 relative rounding + subnormal steps (6e-8 absolute) for the small-weight instances
        _close_per_instance(z.grad, g["d_logits"], 2e-3 if half else GRAD_RTOL, "d_logits per instance", atol=6e-8 if half else 0.0)
    else:
        assert z.grad is None


@pytest.mark.parametrize("name", golden_names("loss"))
def test_golden_weighted_ce(cuda_device, name):
    import deep_staple_b200 as ds

    g = load_golden(name)
    half = g["logits"].dtype != np.float32
    z = torch.from_numpy(g["logits"]).to(cuda_device).requires_grad_(True)
    target = torch.from_numpy(g["target"].astype(np.int64)).to(cuda_device)
    cw = torch.from_numpy(g["cw"]).to(cuda_device)
    scale = float(g["g"])
    ce = ds.weighted_ce(z, target, cw, grad_scaler=scale)
    (ce * scale).backward()  # scaler.scale(ce_loss).backward(), :719
    _close(ce, g["ce_loss"], LOSS_RTOL, "ce_loss")
    _close(z.grad, g["d_logits_ce"], 2e-3 if half else GRAD_RTOL, "d_logits_ce")


def test_backward_without_scale_hint_is_reconciled(cuda_device):
    """No grad_scaler hint: the fused pass assumes g = 1 and backward rescales the stored gradients."""
    import deep_staple_b200 as ds

    g = load_golden("loss__b4_fp16_scale65536")
    dev = cuda_device
    z = torch.from_numpy(g["logits"]).to(dev).requires_grad_(True)
    target = torch.from_numpy(g["target"].astype(np.int64)).to(dev)
    bare = torch.from_numpy(g["emb"]).reshape(-1)[torch.from_numpy(g["idx"])].to(dev).requires_grad_(True)
    fixed = torch.from_numpy(g["fixed"])[torch.from_numpy(g["idx"])].to(dev)
    loss = ds.dp_weighted_loss(z, target, bare, fixed, True)
    (loss * float(g["g"])).backward()
    _close(bare.grad, g["d_bare_weight"], GRAD_RTOL, "d_bare_weight")
    _close(z.grad, g["d_logits"], 4e-3, "d_logits")  # two fp16 roundings on this path


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("C,shape", [(2, (5, 7, 9)), (3, (4, 4, 8)), (5, (3, 5, 7)), (8, (4, 8, 8)), (2, (33, 17)),
                                     (4, (16, 24))])
def test_random_vs_oracle(cuda_device, dtype, C, shape):
    """odd sizes (scalar path), C up to 8 (run-time class count), half logits, 2D slice mode."""
    import deep_staple_b200 as ds

    gen = torch.Generator().manual_seed(1000 + C + len(shape))
    B, M = 3, 10
    logits = (2 * torch.randn((B, C, *shape), generator=gen)).to(dtype)
    target = torch.randint(0, C, (B, *shape), generator=gen)
    emb_w = torch.randn((M, 1), generator=gen)
    idx = torch.randperm(M, generator=gen)[:B]
    fixed = torch.rand(M, generator=gen) * 5 + 1
    cw = torch.rand(C, generator=gen) + 0.5
    ref = ol.run_loss(logits, target, emb_w, idx, fixed, cw, True, 8.0, True)
    dev = cuda_device
    z = logits.to(dev).requires_grad_(True)
    bare = emb_w.reshape(-1)[idx].to(dev).requires_grad_(True)
    out = ds.dp_weighted_loss(z, target.to(dev), bare, fixed[idx].to(dev), True, grad_scaler=8.0,
                              class_weights=cw.to(dev), return_all=True)
    (out.dp_loss * 8.0).backward()
    half = dtype != torch.float32
    _close(out.dp_loss, ref.dp_loss, LOSS_RTOL, "dp_loss")
    _close(out.ce_loss, ref.ce_loss, LOSS_RTOL, "ce_loss")
    _close(out.per_instance_ce, ref.per_inst_ce, LOSS_RTOL, "per_instance_ce")
    assert torch.equal(out.fg_pred.cpu().long(), ref.fg_pred)
    _close(bare.grad, ref.d_bare_weight, GRAD_RTOL, "d_bare_weight")
    _close(z.grad.float(), ref.d_logits.float(), (1e-2 if dtype == torch.bfloat16 else 2e-3) if half else GRAD_RTOL, "d_logits")
    _close_per_instance(z.grad.float(), ref.d_logits.float(), (1e-2 if dtype == torch.bfloat16 else 2e-3) if half else GRAD_RTOL,
                        "d_logits per instance", atol=6e-8 if dtype == torch.float16 else 0.0)
    # hard-Dice statistics of the same pass
    pred = logits.float().argmax(1)
    for c in range(C):
        assert int(out.inst_stats[:, c, 1].sum()) == int((target == c).sum())
        assert int(out.inst_stats[:, c, 2].sum()) == int((pred == c).sum())
        assert int(out.inst_stats[:, c, 3].sum()) == int(((pred == c) & (target == c)).sum())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("B,C,shape", [
    (33, 2, (32, 64)),        # > 32 instances: block-wide epilogue reduction
    (64, 2, (256, 256)),      # cfg4 (2D slice mode): dynamic tail with 8 claim counters x 64 instances
    (70, 2, (16, 64)),        # > 64 instances: data-parameter prologue fallback
    (300, 2, (8, 16)),        # > 256 instances: generic (loop) epilogue
    (2, 2, (96, 128, 128)),   # long instances: static rounds + dynamic tail (LDG shell) / TMA shell (fp16)
    (3, 3, (64, 96, 128)),    # C = 3: TMA shell
])
def test_batch_and_schedule_paths_vs_oracle(cuda_device, dtype, B, C, shape):
    """Every epilogue form (one warp / block / loop), every prologue form (<= 64 instances per warp pass, block
    sum beyond) and both schedules of the streaming loop (static rounds + per-warp claimed tail, TMA tiles),
    each against the oracle -- plus bit-identical statistics and outputs on a second run (the claimed tail
    changes which warp processes which step, never the exact integer sums)."""
    import deep_staple_b200 as ds

    gen = torch.Generator().manual_seed(77 + B)
    M = B + 5
    logits = (2 * torch.randn((B, C, *shape), generator=gen)).to(dtype)
    target = (torch.rand((B, *shape), generator=gen) * C * 0.6).long().clamp_(0, C - 1)
    emb_w = torch.randn((M, 1), generator=gen)
    idx = torch.randperm(M, generator=gen)[:B]
    fixed = torch.rand(M, generator=gen) * 5 + 1
    cw = torch.rand(C, generator=gen) + 0.5
    ref = ol.run_loss(logits, target, emb_w, idx, fixed, cw, True, 64.0, True)
    dev = cuda_device
    outs = []
    for _ in range(2):
        z = logits.to(dev).requires_grad_(True)
        bare = emb_w.reshape(-1)[idx].to(dev).requires_grad_(True)
        out = ds.dp_weighted_loss(z, target.to(dev), bare, fixed[idx].to(dev), True, grad_scaler=64.0,
                                  class_weights=cw.to(dev), return_all=True)
        (out.dp_loss * 64.0).backward()
        outs.append((out, z.grad.clone(), bare.grad.clone()))
    out, dz, dbare = outs[0]
    _close(out.dp_loss, ref.dp_loss, LOSS_RTOL, "dp_loss")
    _close(out.ce_loss, ref.ce_loss, LOSS_RTOL, "ce_loss")
    _close(out.per_instance_ce, ref.per_inst_ce, LOSS_RTOL, "per_instance_ce")
    assert torch.equal(out.fg_pred.cpu().long(), ref.fg_pred)
    _close(dbare, ref.d_bare_weight, GRAD_RTOL, "d_bare_weight")
    # fp16 gradients below 6.1e-5 are subnormal (absolute step 5.96e-8): the long instances land there at scale 64
    _close(dz.float(), ref.d_logits.float(), 2e-3 if dtype == torch.float16 else GRAD_RTOL, "d_logits",
           atol=6e-8 if dtype == torch.float16 else 0.0)
    _close_per_instance(dz.float(), ref.d_logits.float(), 2e-3 if dtype == torch.float16 else GRAD_RTOL, "d_logits per instance",
                        atol=6e-8 if dtype == torch.float16 else 0.0)
    out2, dz2, dbare2 = outs[1]
    assert torch.equal(out.inst_stats, out2.inst_stats) and torch.equal(out.dp_loss, out2.dp_loss)
    assert torch.equal(dz, dz2) and torch.equal(dbare, dbare2)


def test_cfg2_full_size_vs_oracle(cuda_device):
    """BASELINE.json configs[1]: batch 8 x 2 classes x 128^3, risk regularisation, fwd+bwd."""
    import deep_staple_b200 as ds

    gen = torch.Generator().manual_seed(0)
    B, C, S, M = 8, 2, (128, 128, 128), 600
    logits = 2 * torch.randn((B, C, *S), generator=gen)
    target = (torch.rand((B, *S), generator=gen) < 0.05).long()
    emb_w = torch.randn((M, 1), generator=gen)
    idx = torch.randperm(M, generator=gen)[:B]
    gt_num = torch.randint(0, 50000, (M,), generator=gen).float()
    fixed = (gt_num + math.e).log() + math.e
    counts = torch.bincount(target.reshape(-1), minlength=C).float()
    cw = 1 / counts.pow(0.35)
    cw = cw / cw.mean()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref = ol.run_loss(logits, target, emb_w, idx, fixed, cw, True, 65536.0, True)
    dev = cuda_device
    z = logits.to(dev).requires_grad_(True)
    bare = emb_w.reshape(-1)[idx].to(dev).requires_grad_(True)
    out = ds.dp_weighted_loss(z, target.to(dev), bare, fixed[idx].to(dev), True, grad_scaler=65536.0,
                              class_weights=cw.to(dev), return_all=True)
    (out.dp_loss * 65536.0).backward()
    e1 = _close(out.dp_loss, ref.dp_loss, LOSS_RTOL, "dp_loss")
    e2 = _close(out.ce_loss, ref.ce_loss, LOSS_RTOL, "ce_loss")
    e3 = _close(bare.grad, ref.d_bare_weight, GRAD_RTOL, "d_bare_weight")
    e4 = _close(z.grad, ref.d_logits, GRAD_RTOL, "d_logits")
    assert torch.equal(out.fg_pred.cpu().long(), ref.fg_pred)
    print(f"[cfg2] rel err: dp_loss {e1:.2e} ce_loss {e2:.2e} d_bare {e3:.2e} d_logits {e4:.2e}")
    # run-to-run determinism (fixed-order reductions, no float atomics)
    z2 = logits.to(dev).requires_grad_(True)
    out2 = ds.dp_weighted_loss(z2, target.to(dev), bare.detach(), fixed[idx].to(dev), True, grad_scaler=65536.0,
                               class_weights=cw.to(dev), return_all=True)
    assert torch.equal(out2.dp_loss, out.dp_loss) and torch.equal(out2.inst_stats, out.inst_stats)


def test_cfg5_instance_size_vs_oracle(cuda_device):
    """cfg5 (BASELINE.json configs[4]) at its real instance size, 3 classes x 192 x 160 x 160, two instances:
    the TMA-staged shell over 4800 tiles per instance against the oracle, loss 1e-5 / gradients 1e-4 relative."""
    import deep_staple_b200 as ds

    gen = torch.Generator().manual_seed(55)
    B, C, shape, M = 2, 3, (192, 160, 160), 9
    logits = 2 * torch.randn((B, C, *shape), generator=gen)
    target = torch.multinomial(torch.tensor([0.94, 0.05, 0.01]), B * math.prod(shape), replacement=True,
                               generator=gen).reshape(B, *shape)
    emb_w = torch.randn((M, 1), generator=gen)
    idx = torch.randperm(M, generator=gen)[:B]
    fixed = torch.rand(M, generator=gen) * 5 + 1
    cw = torch.tensor([0.4, 1.3, 2.1])
    ref = ol.run_loss(logits, target, emb_w, idx, fixed, cw, True, 65536.0, True)
    dev = cuda_device
    z = logits.to(dev).requires_grad_(True)
    bare = emb_w.reshape(-1)[idx].to(dev).requires_grad_(True)
    out = ds.dp_weighted_loss(z, target.to(dev), bare, fixed[idx].to(dev), True, grad_scaler=65536.0,
                              class_weights=cw.to(dev), return_all=True)
    (out.dp_loss * 65536.0).backward()
    _close(out.dp_loss, ref.dp_loss, LOSS_RTOL, "dp_loss")
    _close(out.ce_loss, ref.ce_loss, LOSS_RTOL, "ce_loss")
    assert torch.equal(out.fg_pred.cpu().long(), ref.fg_pred)
    _close(bare.grad, ref.d_bare_weight, GRAD_RTOL, "d_bare_weight")
    _close(z.grad, ref.d_logits, GRAD_RTOL, "d_logits")


def test_two_pass_backward_matches_fused(cuda_device):
    import deep_staple_b200 as ds

    gen = torch.Generator().manual_seed(9)
    logits = torch.randn((4, 3, 8, 8, 8), generator=gen).to(cuda_device)
    target = torch.randint(0, 3, (4, 8, 8, 8), generator=gen).to(cuda_device)
    bare = torch.randn(4, generator=gen).to(cuda_device)
    z = logits.clone().requires_grad_(True)
    out = ds.dp_weighted_loss(z, target, bare, None, True, grad_scaler=4.0, return_all=True)
    (out.dp_loss * 4.0).backward()
    two = ds.backward_logits(logits, target, out.instance_weight, 4.0)
    assert torch.allclose(two, z.grad, rtol=1e-6, atol=1e-12)


def test_prepass_and_dice_golden(cuda_device):
    import deep_staple_b200 as ds

    for name in golden_names("prepass"):
        g = load_golden(name)
        labels = torch.from_numpy(g["labels"].astype(np.int64)).to(cuda_device)
        cw, fw, counts, gt = ds.prepass_constants(labels, int(g["num_classes"]))
        assert np.array_equal(counts.cpu().numpy(), g["counts"])
        assert np.array_equal(gt.cpu().numpy(), g["gt_num"])
        assert np.allclose(cw.cpu().numpy(), g["class_weights"], rtol=1e-6)
        assert np.allclose(fw.cpu().numpy(), g["fixed_weighting"], rtol=1e-6)
    for name in golden_names("dice"):
        g = load_golden(name)
        for dt in (torch.int64, torch.uint8):
            d = ds.dice_from_label_maps(torch.from_numpy(g["pred"]).to(dt), torch.from_numpy(g["target"]).to(dt),
                                        int(g["num_classes"]), bool(g["nan_for_unlabeled_target"]))
            assert d.device.type == "cpu" and d.dtype == torch.float32
            assert np.allclose(d.numpy(), g["dice"], rtol=1e-6, atol=0, equal_nan=True)


def test_sparse_adam_and_grad_scaler_flow(cuda_device):
    """The reference's optimiser objects keep working around the op: nn.Embedding(sparse=True) +
    SparseAdam + GradScaler (main_deep_staple.py:418, 442-445, 762-773)."""
    import deep_staple_b200 as ds

    torch.manual_seed(0)
    dev = cuda_device
    M, B = 50, 4
    emb = nn.Embedding(M, 1, sparse=True).to(dev)
    torch.nn.init.normal_(emb.weight.data, mean=0.0, std=0.0)
    opt = torch.optim.SparseAdam(emb.parameters(), lr=0.1, betas=(0.9, 0.999), eps=1e-8)
    scaler = torch.amp.GradScaler("cuda")
    logits = torch.randn((B, 2, 8, 8, 8), device=dev, dtype=torch.float16)
    target = (torch.rand((B, 8, 8, 8), device=dev) < 0.3).long()
    idx = torch.tensor([3, 7, 11, 19], device=dev)
    before = emb.weight.detach().clone()
    loss = ds.dp_weighted_loss(logits, target, (emb, idx), None, True, grad_scaler=scaler)
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()
    assert isinstance(loss.item(), float)
    changed = (emb.weight.detach() != before).reshape(-1).nonzero().reshape(-1)
    assert set(changed.tolist()) <= set(idx.tolist()) and len(changed) > 0


def test_invalid_arguments_raise(cuda_device):
    import deep_staple_b200 as ds

    z = torch.zeros((2, 2, 4, 4), device=cuda_device)
    t = torch.zeros((2, 4, 4), dtype=torch.long, device=cuda_device)
    with pytest.raises(ValueError):
        ds.dp_weighted_loss(z, t[:1], torch.zeros(2, device=cuda_device))
    with pytest.raises(ValueError):
        ds.dp_weighted_loss(z, t.float(), torch.zeros(2, device=cuda_device))
    with pytest.raises(ValueError):
        ds.dp_weighted_loss(z, t, torch.zeros(3, device=cuda_device))
    with pytest.raises(ValueError):
        ds.dp_weighted_loss(torch.zeros((2, 9, 4, 4), device=cuda_device), t, torch.zeros(2, device=cuda_device))
    with pytest.raises(RuntimeError):
        ds.dp_weighted_loss(z.cpu(), t.cpu(), torch.zeros(2))


def test_cuda_graph_capture_and_replay(cuda_device):
    """The C-ABI calls allocate nothing and never synchronise: the fused loss + its backward glue can be
    captured in a CUDA graph and replayed on new data (threading / streams)."""
    import deep_staple_b200 as ds
    from deep_staple_b200 import _binding as b

    dev = cuda_device
    lib = b.load()
    gen = torch.Generator().manual_seed(3)
    B, C, S = 4, 2, (16, 16, 16)
    V = 16 ** 3
    z = torch.zeros((B, C, *S), device=dev)
    t = torch.zeros((B, *S), dtype=torch.long, device=dev)
    bare = torch.randn(B, generator=gen).to(dev)
    fixed = (torch.rand(B, generator=gen) + 1).to(dev)
    cw = torch.tensor([0.7, 1.3], device=dev)
    scale = torch.tensor(256.0, device=dev)
    dl = torch.empty_like(z)
    stats = torch.zeros((B, C, 4), dtype=torch.float64, device=dev)
    loss_out = torch.zeros(2, device=dev)
    d_bare = torch.zeros(B, device=dev)
    per_inst = torch.zeros((B, 4), device=dev)
    ws = torch.zeros((int(lib.ds_dp_loss_workspace_bytes(B, C)) + 7) // 8, dtype=torch.int64, device=dev)

    def launch():
        rc = lib.ds_dp_loss_fwd_bwd(z.data_ptr(), b.DS_F32, t.data_ptr(), b.DS_I64, None, B, C, V, bare.data_ptr(),
                                    fixed.data_ptr(), B, 0, cw.data_ptr(), 1, scale.data_ptr(), dl.data_ptr(),
                                    stats.data_ptr(), loss_out.data_ptr(), d_bare.data_ptr(), per_inst.data_ptr(),
                                    ws.data_ptr(), ws.numel() * 8, torch.cuda.current_stream(dev).cuda_stream)
        assert rc == 0

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        launch()  # warm-up outside the capture
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        launch()
    for seed in (10, 11):
        g2 = torch.Generator().manual_seed(seed)
        logits = 2 * torch.randn((B, C, *S), generator=g2)
        target = (torch.rand((B, *S), generator=g2) < 0.3).long()
        z.copy_(logits)
        t.copy_(target)
        graph.replay()
        torch.cuda.synchronize()
        zz = z.clone().requires_grad_(True)
        ref = ds.dp_weighted_loss(zz, t, bare, fixed, True, grad_scaler=256.0, class_weights=cw, return_all=True)
        (ref.dp_loss * 256.0).backward()
        assert torch.equal(loss_out[0], ref.dp_loss.detach()) and torch.equal(stats, ref.inst_stats)
        assert torch.equal(dl, zz.grad)


def test_public_api_fast_paths_are_bit_identical(cuda_device):
    """The lean Python layer (one output arena, no autograd node when nothing requires grad, one reconciling launch in
    backward) must return exactly what the raw C-ABI call returns, in every mode."""
    import deep_staple_b200 as ds
    from deep_staple_b200 import _binding as b

    dev = cuda_device
    gen = torch.Generator().manual_seed(11)
    B, C, S, M = 5, 2, (8, 16, 16), 20
    V = S[0] * S[1] * S[2]
    logits = (2 * torch.randn((B, C, *S), generator=gen)).to(dev)
    target = (torch.rand((B, *S), generator=gen) < 0.2).long().to(dev)
    bare = torch.randn(B, generator=gen).to(dev)
    fixed = (torch.rand(B, generator=gen) * 3 + 1).to(dev)
    cw = torch.tensor([0.6, 1.4], device=dev)
    scale = torch.tensor(512.0, device=dev)
    lib = b.load()
    dz = torch.empty_like(logits)
    stats = torch.empty((B, C, 4), dtype=torch.float64, device=dev)
    loss_out, d_bare, per_inst = torch.empty(2, device=dev), torch.empty(B, device=dev), torch.empty((B, 4), device=dev)
    ws = torch.zeros((int(lib.ds_dp_loss_workspace_bytes(B, C)) + 7) // 8, dtype=torch.int64, device=dev)
    rc = lib.ds_dp_loss_fwd_bwd(logits.data_ptr(), b.DS_F32, target.data_ptr(), b.DS_I64, None, B, C, V, bare.data_ptr(),
                                fixed.data_ptr(), B, 0, cw.data_ptr(), 1, scale.data_ptr(), dz.data_ptr(), stats.data_ptr(),
                                loss_out.data_ptr(), d_bare.data_ptr(), per_inst.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    # no autograd at all
    out = ds.dp_weighted_loss(logits, target, bare, fixed, True, grad_scaler=scale, class_weights=cw, return_all=True)
    assert out.dp_loss.grad_fn is None
    assert torch.equal(out.dp_loss, loss_out[0]) and torch.equal(out.ce_loss, loss_out[1])
    assert torch.equal(out.inst_stats, stats) and torch.equal(out.d_bare_weight, d_bare)
    assert torch.equal(out.per_instance_ce, per_inst[:, 0]) and torch.equal(out.fg_pred, per_inst[:, 1])
    assert torch.equal(out.instance_weight, per_inst[:, 2])
    # autograd, GradScaler flow: gradients are the fused pass's, bit for bit
    z = logits.clone().requires_grad_(True)
    p = bare.clone().requires_grad_(True)
    loss = ds.dp_weighted_loss(z, target, p, fixed, True, grad_scaler=scale, class_weights=cw)
    (loss * scale).backward(retain_graph=True)
    assert torch.equal(loss.detach(), loss_out[0]) and torch.equal(z.grad, dz) and torch.equal(p.grad, d_bare)
    # a second backward with ANOTHER incoming gradient (retain_graph): reconciled once, not twice
    z.grad, p.grad = None, None
    (loss * (scale * 0.25)).backward()
    _close(z.grad, dz * 0.25, 1e-6, "second backward, rescaled dlogits")
    _close(p.grad, d_bare * 0.25, 1e-6, "second backward, rescaled d_bare")
    # half precision WITHOUT the scale hint: gradients are recomputed in backward, not stored as fp16 subnormals
    zh = logits.half().requires_grad_(True)
    ph = bare.clone().requires_grad_(True)
    lh = ds.dp_weighted_loss(zh, target, ph, fixed, True)
    (lh * 65536.0).backward()
    zr = logits.half().requires_grad_(True)
    lr_ = ds.dp_weighted_loss(zr, target, bare.clone().requires_grad_(True), fixed, True, grad_scaler=65536.0)
    (lr_ * 65536.0).backward()
    assert torch.equal(zh.grad, zr.grad), "two-pass backward must equal the fused, pre-scaled gradient"
    # class-weighted CE: same contract
    zc = logits.half().requires_grad_(True)
    ce = ds.weighted_ce(zc, target, cw)
    (ce * 65536.0).backward()
    zc2 = logits.half().requires_grad_(True)
    ce2 = ds.weighted_ce(zc2, target, cw, grad_scaler=65536.0)
    (ce2 * 65536.0).backward()
    assert torch.equal(ce.detach(), ce2.detach()) and torch.equal(zc.grad, zc2.grad)
