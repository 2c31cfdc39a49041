"""GPU STAPLE (EM) baseline (ds_staple) against oracle/staple.py -- the fp64 restatement of the published algorithm in
the form ITK's STAPLEImageFilter runs it (consensus.ipynb:112-124 calls SimpleITK, which is third-party and not
installable here: PARITY UNPINNED against ITK itself, stated in DESIGN.md)."""
import numpy as np
import pytest
import torch

from oracle import staple as os_

pytestmark = pytest.mark.gpu


def _raters(n, shape, seed, quality=(0.7, 0.98)):
    """A hidden truth (one ellipsoid) seen by n raters of different sensitivity / specificity."""
    gen = torch.Generator().manual_seed(seed)
    zz, yy, xx = torch.meshgrid(*[torch.arange(s, dtype=torch.float32) for s in shape], indexing="ij")
    c = torch.tensor(shape, dtype=torch.float32) / 2
    truth = (((zz - c[0]) / (shape[0] / 3)) ** 2 + ((yy - c[1]) / (shape[1] / 4)) ** 2 + ((xx - c[2]) / (shape[2] / 3)) ** 2) <= 1
    out = []
    for _ in range(n):
        p = quality[0] + (quality[1] - quality[0]) * float(torch.rand((), generator=gen))
        q = 0.9 + 0.099 * float(torch.rand((), generator=gen))
        u = torch.rand(shape, generator=gen)
        out.append(torch.where(truth, u < p, u > q).long())
    return out, truth


@pytest.mark.parametrize("n,shape,dtype", [(10, (16, 20, 24), torch.int64), (30, (12, 16, 32), torch.uint8), (3, (8, 8, 8), torch.int64),
                                           (40, (8, 12, 16), torch.uint8), (17, (7, 5, 3), torch.int64)])
def test_staple_matches_the_fp64_restatement(cuda_device, n, shape, dtype):
    import deep_staple_b200 as ds

    lbls, truth = _raters(n, shape, 5 + n)
    w, p, q, it = os_.staple(np.stack([l.numpy().reshape(-1) for l in lbls]), 1.0, 200)
    cons, sens, spec = ds.calc_staple_consensus([l.to(dtype).to(cuda_device) for l in lbls])
    assert cons.dtype == torch.int64 and tuple(cons.shape) == shape and cons.device.type == "cuda"
    ref = torch.from_numpy((w > 0.5).reshape(shape))
    knife = torch.from_numpy((np.abs(w - 0.5) < 1e-4).reshape(shape))  # posterior within 1e-4 of the threshold
    assert not bool(((cons.cpu().bool() != ref) & ~knife).any())
    assert np.allclose(sens, p, rtol=0, atol=2e-5) and np.allclose(spec, q, rtol=0, atol=2e-5), (np.abs(np.array(sens) - p).max(),
                                                                                               np.abs(np.array(spec) - q).max())
    # and it does what STAPLE is for: the estimate is closer to the hidden truth than a plain majority vote of noisy raters
    maj = torch.stack(lbls).float().mean(0) > 0.5
    dice = lambda a, b: 2.0 * float((a & b).sum()) / float(a.sum() + b.sum())
    assert dice(cons.cpu().bool(), truth) >= dice(maj, truth) - 1e-9


def test_staple_batched_probability_and_iterations(cuda_device):
    import deep_staple_b200 as ds

    T, n, shape = 3, 12, (8, 16, 16)
    stacks, refs = [], []
    for t in range(T):
        lbls, _ = _raters(n, shape, 100 + t)
        stacks.append(torch.stack(lbls))
        refs.append(os_.staple(np.stack([l.numpy().reshape(-1) for l in lbls]), 1.0, 200))
    labels = torch.stack(stacks).to(torch.uint8).to(cuda_device)
    out, sens, spec, iters, prob = ds.staple_batched(labels, return_probability=True)
    for t in range(T):
        w, p, q, it = refs[t]
        assert abs(int(iters[t]) - it) <= 2, (int(iters[t]), it)  # fp32 partial sums can move the 1e-7 stopping test by an iteration
        assert np.abs(prob[t].cpu().numpy().reshape(-1) - w).max() < 1e-4
        assert np.allclose(sens[t].cpu().numpy(), p, atol=2e-5) and np.allclose(spec[t].cpu().numpy(), q, atol=2e-5)
    # sparse COO input (what the snapshot stores) and the maximum-iterations cut
    sp = [l.to_sparse() for l in stacks[0]]
    cons, s1, _ = ds.calc_staple_consensus(sp, max_iterations=3)
    w3, p3, _, it3 = os_.staple(np.stack([l.numpy().reshape(-1) for l in stacks[0]]), 1.0, 3)
    assert it3 == 3 and cons.device.type == "cpu"
    assert np.allclose(s1, p3, atol=2e-5)
    assert int((cons.bool() != torch.from_numpy((w3 > 0.5).reshape(shape))).sum()) == 0
    with pytest.raises(ValueError):
        ds.staple_batched(torch.zeros((1, 65, 4, 4, 4), dtype=torch.uint8, device=cuda_device))


def test_snapshot_driver_with_staple_keys(cuda_device):
    """consensus.ipynb:393-418: with_staple adds staple_consensus / staple_sensitivity / staple_consensus_oracle_dice."""
    import deep_staple_b200 as ds

    torch.manual_seed(0)
    shape = (8, 8, 16)
    ids = ["100l:m200l", "100l:m201r", "100l:m202l", "101r:m200l", "101r:m201r", "101r:m202l"]
    mod = torch.stack([(torch.rand(shape) < 0.4).long().to_sparse() for _ in ids])
    lab = torch.stack([(torch.rand(shape) < 0.4).long().to_sparse() for _ in ids])
    pred = torch.stack([(torch.rand(1, *shape) < 0.4).long().to_sparse() for _ in ids])
    snap = {"d_ids": tuple(ids), "data_parameters": torch.randn(len(ids), 1), "labels": lab, "modified_labels": mod,
            "train_predictions": pred}
    out = ds.consensus_from_snapshot(snap, with_staple=True)
    for f_id in ("100l", "101r"):
        d = out[f_id]
        assert d["staple_consensus"].is_sparse and d["staple_consensus_oracle_dice"].shape == (1, 2)
        for m in ("200l", "201r", "202l"):
            assert 0.0 <= d[m]["staple_sensitivity"] <= 1.0
        ref, _, _ = os_.calc_staple_consensus([d[m]["warped_label"] for m in ("200l", "201r", "202l")])
        assert int((d["staple_consensus"].to_dense() != ref).sum()) == 0
