"""GPU parity on NON-FINITE logits: the reference's AMP flow needs NaN / Inf to reach the loss and the gradients so
that ``scaler_dp.step(optimizer_dp)`` skips the step and ``update()`` backs the scale off
(main_deep_staple.py:762-768).  Fixtures ``tests/golden/nonfinite__*.npz`` hold the outputs of the reference's own
lines on +Inf / -Inf / NaN / overflowing-fp16 logits, in-line and out-of-line; the CUDA path must reproduce the
NaN / Inf pattern exactly and every finite value within the usual tolerances."""
import numpy as np
import pytest
import torch
from torch import nn

from conftest import golden_names, load_golden
from oracle import dp_loss as ol

pytestmark = pytest.mark.gpu


def _same_pattern(a, b, what, rtol, atol=0.0):
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1).numpy()
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN pattern differs: {a} vs {b}" if a.size < 10 else f"{what}: NaN pattern differs"
    assert np.array_equal(np.isposinf(a), np.isposinf(b)) and np.array_equal(np.isneginf(a), np.isneginf(b)), f"{what}: Inf pattern differs"
    fin = np.isfinite(b)
    if fin.any():
        scale = np.abs(b[fin]).max()
        err = np.abs(a[fin] - b[fin]).max()
        assert err <= rtol * scale + atol, f"{what}: err {err:.3e} vs scale {scale:.3e}"


def _run(g, dev, label_dtype=torch.int64):
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
    out = ds.dp_weighted_loss(z, target, emb(idx).squeeze(), fixed, bool(g["risk"]), grad_scaler=scale, class_weights=cw,
                              return_all=True)
    (out.dp_loss * scale).backward()
    return out, z, emb, idx


@pytest.mark.parametrize("name", golden_names("nonfinite"))
@pytest.mark.parametrize("label_dtype", [torch.int64, torch.uint8])
def test_nonfinite_dp_loss_matches_reference(cuda_device, name, label_dtype):
    g = load_golden(name)
    half = g["logits"].dtype != np.float32
    out, z, emb, idx = _run(g, cuda_device, label_dtype)
    _same_pattern(out.dp_loss, g["dp_loss"], "dp_loss", 1e-5)
    _same_pattern(out.ce_loss, g["ce_loss"], "ce_loss", 1e-5)
    _same_pattern(out.per_instance_ce, g["per_inst_ce"], "per_instance_ce", 1e-5)
    assert np.array_equal(out.fg_pred.cpu().numpy().astype(np.int64), g["fg_pred"]), "fg_pred (torch.argmax semantics on NaN)"
    rows = emb.weight.grad.coalesce().to_dense().reshape(-1)[idx]
    _same_pattern(rows, g["d_bare_weight"], "d_bare_weight", 1e-4, atol=1e-9 * float(g["g"]))
    _same_pattern(out.d_bare_weight, g["d_bare_weight"], "d_bare_weight (fused output)", 1e-4, atol=1e-9 * float(g["g"]))
    if bool(g["need_dlogits"]):
        _same_pattern(z.grad, g["d_logits"], "d_logits", 2e-3 if half else 1e-4)
    # the statistics of the same pass: a NaN / Inf loss sum sits in the class of the offending voxel's label only
    ref_nll = np.zeros((g["logits"].shape[0], g["logits"].shape[1]))
    zf = torch.from_numpy(g["logits"]).float()
    t = torch.from_numpy(g["target"].astype(np.int64))
    nll = nn.CrossEntropyLoss(reduction="none")(zf, t).double()
    for c in range(g["logits"].shape[1]):
        ref_nll[:, c] = torch.where(t == c, nll, torch.zeros_like(nll)).flatten(1).sum(1).numpy()
    _same_pattern(out.inst_stats[:, :, 0], ref_nll, "per-class loss sums", 1e-5)


@pytest.mark.parametrize("name", golden_names("nonfinite"))
def test_nonfinite_weighted_ce_matches_reference(cuda_device, name):
    import deep_staple_b200 as ds

    g = load_golden(name)
    half = g["logits"].dtype != np.float32
    z = torch.from_numpy(g["logits"]).to(cuda_device).requires_grad_(True)
    target = torch.from_numpy(g["target"].astype(np.int64)).to(cuda_device)
    scale = float(g["g"])
    ce = ds.weighted_ce(z, target, torch.from_numpy(g["cw"]).to(cuda_device), grad_scaler=scale)
    (ce * scale).backward()
    _same_pattern(ce, g["ce_loss"], "ce_loss", 1e-5)
    _same_pattern(z.grad, g["d_logits_ce"], "d_logits_ce", 2e-3 if half else 1e-4)


@pytest.mark.parametrize("kind", ["posinf", "neginf_target", "nan", "fp16_overflow"])
@pytest.mark.parametrize("C,shape,dtype", [(2, (16, 32, 32), torch.float32), (3, (16, 32, 32), torch.float32),
                                           (2, (16, 32, 32), torch.float16), (3, (16, 32, 32), torch.float16),
                                           (3, (16, 32, 32), torch.bfloat16), (4, (8, 32, 32), torch.float32),
                                           (5, (5, 7, 9), torch.float32), (2, (64, 64), torch.float16)])
def test_nonfinite_vs_oracle_every_kernel_shell(cuda_device, kind, C, shape, dtype):
    """Larger shapes (vector path, TMA shell for C >= 3 fp32 and half C = 2, dynamic tail) against the CPU oracle:
    one offending voxel somewhere inside instance 1."""
    import deep_staple_b200 as ds

    if kind == "fp16_overflow" and dtype != torch.float16:
        pytest.skip("fp16 only")
    gen = torch.Generator().manual_seed(7 + C)
    B, M = 3, 9
    logits = (2 * torch.randn((B, C, *shape), generator=gen)).to(dtype)
    target = torch.randint(0, C, (B, *shape), generator=gen)
    flat, tflat = logits.view(B, C, -1), target.view(B, -1)
    v = flat.shape[2] // 2 + 3
    tflat[1, v] = C - 1
    if kind == "posinf":
        flat[1, 0, v] = float("inf")
    elif kind == "neginf_target":
        flat[1, C - 1, v] = float("-inf")
    elif kind == "nan":
        flat[1, C - 1, v] = float("nan")
    else:
        flat[1, C - 1, v] = torch.tensor(70000.0).to(torch.float16)  # inf in fp16
    emb_w = torch.randn((M, 1), generator=gen)
    idx = torch.randperm(M, generator=gen)[:B]
    fixed = torch.rand(M, generator=gen) * 5 + 1
    cw = torch.rand(C, generator=gen) + 0.5
    for need_dlogits in (True, False):
        ref = ol.run_loss(logits, target, emb_w, idx, fixed, cw, True, 256.0, need_dlogits)
        z = logits.to(cuda_device).requires_grad_(need_dlogits)
        bare = emb_w.reshape(-1)[idx].to(cuda_device).requires_grad_(True)
        out = ds.dp_weighted_loss(z, target.to(cuda_device), bare, fixed[idx].to(cuda_device), True, grad_scaler=256.0,
                                  class_weights=cw.to(cuda_device), return_all=True)
        (out.dp_loss * 256.0).backward()
        _same_pattern(out.dp_loss, ref.dp_loss, "dp_loss", 1e-5)
        _same_pattern(out.ce_loss, ref.ce_loss, "ce_loss", 1e-5)
        _same_pattern(out.per_instance_ce, ref.per_inst_ce, "per_instance_ce", 1e-5)
        assert torch.equal(out.fg_pred.cpu().long(), ref.fg_pred)
        _same_pattern(bare.grad, ref.d_bare_weight, "d_bare_weight", 1e-4)
        if need_dlogits:
            tol = 1e-4 if dtype == torch.float32 else (1e-2 if dtype == torch.bfloat16 else 2e-3)
            _same_pattern(z.grad.float(), ref.d_logits.float(), "d_logits", tol)
    # and the workspace is clean again: the next (finite) launch on the same stream is unaffected
    logits2 = torch.nan_to_num(logits.float(), nan=0.0, posinf=1.0, neginf=-1.0).to(dtype)
    ref2 = ol.run_loss(logits2, target, emb_w, idx, fixed, cw, True, 256.0, False)
    out2 = ds.dp_weighted_loss(logits2.to(cuda_device), target.to(cuda_device), emb_w.reshape(-1)[idx].to(cuda_device),
                               fixed[idx].to(cuda_device), True, return_all=True)
    _same_pattern(out2.dp_loss, ref2.dp_loss, "dp_loss after a non-finite launch", 1e-5)


@pytest.mark.parametrize("fused_step", [False, True])
@pytest.mark.parametrize("bad", [float("inf"), float("nan"), float("-inf")])
def test_grad_scaler_skips_the_step_and_backs_off(cuda_device, fused_step, bad):
    """main_deep_staple.py:762-768 end to end: scaler_dp.scale(dp_loss).backward(); scaler_dp.step(optimizer_dp);
    scaler_dp.update() -- a non-finite logit must leave the data parameters and the Adam moments untouched and halve
    the scale, through torch's own SparseAdam / GradScaler and through the fused ds.dp_optimizer_step."""
    import deep_staple_b200 as ds

    dev = cuda_device
    gen = torch.Generator().manual_seed(3)
    B, C, shape, M = 4, 2, (8, 16, 16), 12
    logits = 2 * torch.randn((B, C, *shape), generator=gen)
    target = (torch.rand((B, *shape), generator=gen) < 0.3).long()
    target.view(B, -1)[2, 5] = 1
    idx = torch.randperm(M, generator=gen)[:B].to(dev)
    fixed = (torch.rand(M, generator=gen) * 3 + 1).to(dev)
    emb = nn.Embedding(M, 1, sparse=True).to(dev)
    torch.nn.init.normal_(emb.weight)
    opt = torch.optim.SparseAdam(emb.parameters(), lr=0.1, betas=(0.9, 0.999), eps=1e-8)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)

    def step(z):
        opt.zero_grad()
        out = ds.dp_weighted_loss(z.to(dev), target.to(dev), emb(idx).squeeze(), fixed[idx], True, grad_scaler=scaler,
                                  return_all=True)
        if fused_step:
            scaler.scale(out.dp_loss)  # creates the scale tensor like the reference's scaler_dp.scale(dp_loss)
            applied = ds.dp_optimizer_step(emb, opt, idx, out.d_bare_weight, scaler)
        else:
            scaler.scale(out.dp_loss).backward()
            scaler.step(opt)
            applied = None
        scaler.update()
        return out, applied

    out, applied = step(logits)  # a healthy step first: moments exist, weights move
    assert torch.isfinite(out.dp_loss) and applied in (None, True)
    w0 = emb.weight.detach().clone()
    m0 = opt.state[emb.weight]["exp_avg"].clone()
    v0 = opt.state[emb.weight]["exp_avg_sq"].clone()
    assert scaler.get_scale() == 1024.0
    poisoned = logits.clone()
    poisoned.view(B, C, -1)[2, 1, 5] = bad  # -inf on the target class -> +inf loss, NaN / -inf gradients
    out, applied = step(poisoned)
    assert not torch.isfinite(out.dp_loss)
    assert applied in (None, False)
    assert torch.equal(emb.weight.detach(), w0), "the data parameters moved on a non-finite step"
    assert torch.equal(opt.state[emb.weight]["exp_avg"], m0) and torch.equal(opt.state[emb.weight]["exp_avg_sq"], v0)
    assert scaler.get_scale() == 512.0, "GradScaler.update() must back the scale off"
    out, applied = step(logits)  # and training resumes
    assert torch.isfinite(out.dp_loss) and not torch.equal(emb.weight.detach(), w0)
