"""Property tests (hypothesis) over shapes / dtypes / class counts / candidate counts -- SURVEY.md section 4, tier 3:
whatever the shape, the CUDA path agrees with the CPU oracle within BASELINE.json's tolerances.  Shapes are drawn small
so that the oracle takes milliseconds; the parametrised tests cover the full-size configurations."""
import math

import numpy as np
import pytest
import torch

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

from oracle import consensus as oc  # noqa: E402
from oracle import dp_loss as ol  # noqa: E402

pytestmark = pytest.mark.gpu
SETTINGS = dict(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow],
                derandomize=True)

spatial = st.one_of(
    st.tuples(st.integers(1, 9), st.integers(1, 9), st.integers(1, 33)),  # 3D, odd sizes: scalar and vector paths
    st.tuples(st.integers(1, 40), st.integers(1, 40)),                    # 2D slice mode
    st.tuples(st.integers(1, 6), st.integers(1, 6), st.sampled_from([4, 8, 16, 32, 64])),
)


def _rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu().reshape(-1), torch.as_tensor(b).detach().double().cpu().reshape(-1)
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


@settings(**SETTINGS)
@given(B=st.integers(1, 5), C=st.integers(2, 8), shape=spatial,
       dtype=st.sampled_from([torch.float32, torch.float16, torch.bfloat16]),
       label_dtype=st.sampled_from([torch.int64, torch.uint8]), risk=st.booleans(), use_fixed=st.booleans(),
       inline=st.booleans(), seed=st.integers(0, 10_000))
def test_dp_loss_any_shape_matches_oracle(cuda_device, B, C, shape, dtype, label_dtype, risk, use_fixed, inline, seed):
    import deep_staple_b200 as ds

    gen = torch.Generator().manual_seed(seed)
    M = B + 3
    logits = (2 * torch.randn((B, C, *shape), generator=gen)).to(dtype)
    target = torch.randint(0, C, (B, *shape), generator=gen)
    emb_w = torch.randn((M, 1), generator=gen)
    idx = torch.randperm(M, generator=gen)[:B]
    fixed = torch.rand(M, generator=gen) * 5 + 1 if use_fixed else None
    cw = torch.rand(C, generator=gen) + 0.5
    g = 128.0
    ref = ol.run_loss(logits, target, emb_w, idx, fixed, cw, risk, g, inline)
    dev = cuda_device
    z = logits.to(dev).requires_grad_(inline)
    bare = emb_w.reshape(-1)[idx].to(dev).requires_grad_(True)
    out = ds.dp_weighted_loss(z, target.to(label_dtype).to(dev), bare.squeeze() if B == 1 else bare,
                              fixed[idx].to(dev) if use_fixed else None, risk, grad_scaler=g, class_weights=cw.to(dev),
                              return_all=True)
    (out.dp_loss * g).backward()
    scale = max(abs(float(ref.dp_loss)), 1e-3)  # the risk term can cancel the loss: absolute floor
    assert abs(float(out.dp_loss) - float(ref.dp_loss)) <= 1e-5 * scale + 1e-6
    assert _rel(out.ce_loss, ref.ce_loss) < 1e-5
    assert _rel(out.per_instance_ce, ref.per_inst_ce) < 1e-5
    assert torch.equal(out.fg_pred.cpu().long(), ref.fg_pred)
    gscale = float(ref.d_bare_weight.abs().max())
    assert float((bare.grad.cpu().double().reshape(-1) - ref.d_bare_weight.double()).abs().max()) <= 1e-4 * gscale + 1e-6 * g
    if inline:
        tol = 1e-4 if dtype == torch.float32 else (1e-2 if dtype == torch.bfloat16 else 2e-3)
        a, b = z.grad.float().cpu().double().flatten(1), ref.d_logits.float().double().flatten(1)
        err, own = (a - b).abs().amax(1), b.abs().amax(1)
        assert bool((err <= tol * own + 1e-7).all()), (err.tolist(), own.tolist())  # per instance, against its own scale
    else:
        assert z.grad is None


@settings(**SETTINGS)
@given(N=st.sampled_from([1, 2, 3, 10, 16, 17, 30, 33]), shape=spatial, dtype=st.sampled_from([torch.int64, torch.uint8]),
       sparse=st.booleans(), p=st.floats(0.05, 0.6), spread=st.sampled_from([0.0, 0.3, 2.0]), seed=st.integers(0, 10_000))
def test_consensus_any_shape_bit_exact_outside_near_ties(cuda_device, N, shape, dtype, sparse, p, spread, seed):
    import deep_staple_b200 as ds

    if len(shape) == 2:
        shape = (1, *shape)
    gen = torch.Generator().manual_seed(seed)
    cpu_lbls = [(torch.rand(shape, generator=gen) < p).long() for _ in range(N)]
    dps = [float(x) for x in spread * torch.randn(N, generator=gen)]
    ref = oc.calc_dp_consensus(cpu_lbls, dps)
    if sparse:
        lbls = [l.to_sparse() for l in cpu_lbls]  # CPU sparse COO, what the snapshot holds
    else:
        lbls = [l.to(dtype).to(cuda_device) for l in cpu_lbls]
    out, ties = ds.calc_dp_consensus(lbls, dps, return_near_ties=True)
    near = oc.near_tie_mask(cpu_lbls, dps)
    assert out.dtype == torch.int64 and tuple(out.shape) == tuple(ref.shape)
    assert not bool(((out.cpu() != ref) & ~near).any())
    assert abs(int(ties) - int(near.sum())) <= max(2, int(near.sum()) // 20)
