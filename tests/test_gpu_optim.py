"""Fused data-parameter optimiser step (ds_sparse_adam_step) vs torch's own GradScaler.step(SparseAdam)
(main_deep_staple.py:764-773): same weights, same optimiser state, same scaler behaviour."""
import copy

import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def _setup(dev, M=40, init=0.0):
    emb = nn.Embedding(M, 1, sparse=True).to(dev)
    torch.nn.init.normal_(emb.weight.data, mean=init, std=0.0)  # main_deep_staple.py:438
    opt = torch.optim.SparseAdam(emb.parameters(), lr=0.1, betas=(0.9, 0.999), eps=1e-8)  # :442-444
    scaler = torch.amp.GradScaler("cuda")  # :445
    return emb, opt, scaler


def test_fused_step_matches_gradscaler_plus_sparse_adam(cuda_device):
    import deep_staple_b200 as ds

    dev = cuda_device
    torch.manual_seed(0)
    emb_r, opt_r, sc_r = _setup(dev)
    emb_f, opt_f, sc_f = _setup(dev)
    B, C, S = 6, 2, (8, 8, 16)
    for step in range(5):
        logits = torch.randn((B, C, *S), device=dev, dtype=torch.float16)
        target = (torch.rand((B, *S), device=dev) < 0.3).long()
        idx = torch.randperm(40, device=dev)[:B]
        fixed = torch.rand(B, device=dev) + 1
        # reference flow: autograd -> sparse grad -> GradScaler.step(SparseAdam) -> update   (:762-768)
        loss_r = ds.dp_weighted_loss(logits, target, (emb_r, idx), fixed, True, grad_scaler=sc_r)
        sc_r.scale(loss_r).backward()
        sc_r.step(opt_r)
        sc_r.update()
        opt_r.zero_grad(set_to_none=True)
        # fused flow: no backward at all, one optimiser kernel
        with torch.no_grad():
            bare = emb_f(idx).squeeze()
        sc_f.scale(torch.zeros((), device=dev))  # what scaler.scale(loss) does first: create the scale tensor
        out = ds.dp_weighted_loss(logits, target, bare, fixed, True, grad_scaler=sc_f, return_all=True)
        applied = ds.dp_optimizer_step(emb_f, opt_f, idx, out.d_bare_weight, sc_f)
        sc_f.update()
        assert applied is True
        assert torch.equal(loss_r.detach(), out.dp_loss.detach())
        assert torch.allclose(emb_f.weight, emb_r.weight, rtol=1e-6, atol=1e-9), f"step {step}"
    st_r, st_f = opt_r.state_dict()["state"][0], opt_f.state_dict()["state"][0]
    assert st_r["step"] == st_f["step"] == 5
    assert torch.allclose(st_f["exp_avg"], st_r["exp_avg"], rtol=1e-6, atol=1e-12)
    assert torch.allclose(st_f["exp_avg_sq"], st_r["exp_avg_sq"], rtol=1e-6, atol=1e-16)
    assert sc_f.get_scale() == sc_r.get_scale()
    # the state written by the fused step loads into a fresh torch SparseAdam (checkpoint compatibility)
    emb_n, opt_n, _ = _setup(dev)
    opt_n.load_state_dict(copy.deepcopy(opt_f.state_dict()))
    assert opt_n.state_dict()["state"][0]["step"] == 5


def test_non_finite_gradient_skips_the_step_and_backs_off_the_scale(cuda_device):
    import deep_staple_b200 as ds

    dev = cuda_device
    emb, opt, sc = _setup(dev, init=0.3)
    sc.scale(torch.zeros((), device=dev))
    before = emb.weight.detach().clone()
    idx = torch.tensor([1, 5, 7], device=dev)
    grad = torch.tensor([1.0, float("inf"), 2.0], device=dev)
    applied = ds.dp_optimizer_step(emb, opt, idx, grad, sc)
    scale0 = sc.get_scale()
    sc.update()
    assert applied is False
    assert torch.equal(emb.weight.detach(), before)
    assert opt.state[emb.weight]["step"] == 0
    assert sc.get_scale() == scale0 * 0.5  # GradScaler backoff_factor


def test_without_scaler_and_without_sync(cuda_device):
    import deep_staple_b200 as ds

    dev = cuda_device
    emb, opt, _ = _setup(dev)
    emb_r, opt_r, _ = _setup(dev)
    idx = torch.tensor([0, 3, 9, 11], device=dev)
    grad = torch.tensor([0.5, -0.25, 1.5, -2.0], device=dev)
    for _ in range(3):
        assert ds.dp_optimizer_step(emb, opt, idx, grad, None, sync=False) is None
        emb_r.weight.grad = torch.sparse_coo_tensor(idx.unsqueeze(0), grad.unsqueeze(1), emb_r.weight.shape)
        opt_r.step()
    assert torch.allclose(emb.weight, emb_r.weight, rtol=1e-6, atol=1e-9)
    assert opt.state[emb.weight]["step"] == 3
    with pytest.raises(ValueError):
        ds.dp_optimizer_step(emb, opt, idx, grad[:2], None)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_loss_section_matches_the_reference_flow(cuda_device, dtype):
    """ds.LossSection (two launches, no autograd) against the reference's own sequence (:738-768) run with torch's
    objects: dp_weighted_loss -> scaler.scale(loss).backward() -> scaler.step(SparseAdam) -> scaler.update(),
    including an overflow step in the middle (step skipped, scale halved) and a scale growth (growth_interval = 3)."""
    import deep_staple_b200 as ds

    dev = cuda_device
    torch.manual_seed(1)
    M, B, C, S = 40, 6, 2, (8, 8, 16)

    def setup():
        emb = nn.Embedding(M, 1, sparse=True).to(dev)
        torch.nn.init.normal_(emb.weight.data, mean=0.0, std=0.0)
        opt = torch.optim.SparseAdam(emb.parameters(), lr=0.1, betas=(0.9, 0.999), eps=1e-8)
        return emb, opt, torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=3)

    emb_r, opt_r, sc_r = setup()
    emb_f, opt_f, sc_f = setup()
    fixed_table = torch.rand(M, device=dev) * 3 + 1
    cw = torch.tensor([0.7, 1.3], device=dev)
    section = ds.LossSection(emb_f, opt_f, sc_f, fixed_weighting=fixed_table, class_weights=cw)
    for step in range(9):
        logits = (2 * torch.randn((B, C, *S), device=dev)).to(dtype)
        if step == 4:
            logits.view(B, C, -1)[3, 1, 7] = float("inf")  # overflow: the step must be skipped by both flows
        target = (torch.rand((B, *S), device=dev) < 0.3).long()
        idx = torch.randperm(M, device=dev)[:B]
        loss_r = ds.dp_weighted_loss(logits, target, (emb_r, idx), fixed_table[idx], True, grad_scaler=sc_r)
        sc_r.scale(loss_r).backward()
        sc_r.step(opt_r)
        sc_r.update()
        opt_r.zero_grad(set_to_none=True)
        out = section(logits, target, idx, return_all=True)
        assert torch.equal(loss_r.detach(), out.dp_loss, ) or (torch.isnan(loss_r) and torch.isnan(out.dp_loss)), f"step {step}"
        assert torch.allclose(emb_f.weight, emb_r.weight, rtol=1e-6, atol=1e-9), f"step {step}"
        assert sc_f.get_scale() == sc_r.get_scale(), f"step {step}: scale {sc_f.get_scale()} vs {sc_r.get_scale()}"
        assert int(sc_f._growth_tracker) == int(sc_r._growth_tracker)
    st_r, st_f = opt_r.state_dict()["state"][0], opt_f.state_dict()["state"][0]  # state_dict() syncs the device counter
    assert st_r["step"] == st_f["step"] == 8  # one of the nine steps was skipped
    assert torch.allclose(st_f["exp_avg"], st_r["exp_avg"], rtol=1e-6, atol=1e-12)
    assert torch.allclose(st_f["exp_avg_sq"], st_r["exp_avg_sq"], rtol=1e-6, atol=1e-16)
    same_bits = float((emb_f.weight == emb_r.weight).float().mean())
    print(f"LossSection vs torch flow after 9 steps: {same_bits:.3f} of the data parameters bit-identical")
    assert out.ce_loss.isfinite() and out.inst_stats.shape == (B, C, 4)


def test_loss_section_is_cuda_graph_capturable(cuda_device):
    """Nothing in the section depends on host state: three steps captured once, replayed, equal to eager steps."""
    import deep_staple_b200 as ds

    dev = cuda_device
    torch.manual_seed(2)
    M, B, C, S = 30, 4, 2, (8, 16, 16)

    def setup():
        emb = nn.Embedding(M, 1, sparse=True).to(dev)
        torch.nn.init.normal_(emb.weight.data, mean=0.2, std=0.0)
        opt = torch.optim.SparseAdam(emb.parameters(), lr=0.1)
        return emb, opt, torch.amp.GradScaler("cuda", init_scale=256.0)

    logits = [2 * torch.randn((B, C, *S), device=dev) for _ in range(3)]
    target = [(torch.rand((B, *S), device=dev) < 0.3).long() for _ in range(3)]
    idx = [torch.randperm(M, device=dev)[:B] for _ in range(3)]
    emb_e, opt_e, sc_e = setup()
    eager = ds.LossSection(emb_e, opt_e, sc_e)
    losses_e = [eager(logits[k], target[k], idx[k]).clone() for k in range(3)]
    emb_g, opt_g, sc_g = setup()
    graphed = ds.LossSection(emb_g, opt_g, sc_g)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        graphed(logits[0], target[0], idx[0])  # warm-up on the side stream: workspace for that stream exists
        emb_g.weight.data.fill_(0.2), opt_g.state[emb_g.weight]["exp_avg"].zero_(), opt_g.state[emb_g.weight]["exp_avg_sq"].zero_()
        graphed.step_count.zero_(), sc_g._scale.fill_(256.0), sc_g._growth_tracker.zero_()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            losses_g = [graphed(logits[k], target[k], idx[k]) for k in range(3)]
    g.replay()
    torch.cuda.synchronize()
    for k in range(3):
        assert torch.equal(losses_g[k], losses_e[k]), k
    assert torch.equal(emb_g.weight, emb_e.weight)
    assert graphed.sync() == eager.sync() == 3
