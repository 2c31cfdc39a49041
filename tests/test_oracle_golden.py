"""The oracle restatement against the golden vectors produced by the reference's own code
(tests/golden/make_golden.py), plus a drift guard that re-runs the reference lines when the
reference tree is mounted.  CPU only."""
import math

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from oracle import consensus as oc
from oracle import dp_loss as ol
from oracle import metrics as om
from oracle import reference_lines as ref


@pytest.mark.parametrize("name", golden_names("consensus"))
def test_consensus_oracle_matches_reference_output(name):
    g = load_golden(name)
    lbls = [torch.from_numpy(l.astype(np.int64)) for l in g["labels"]]
    if bool(g["sparse"]):
        lbls = [l.to_sparse() for l in lbls]
    dps = [torch.tensor([d], dtype=torch.float32) for d in g["dps"]]
    out = oc.calc_dp_consensus(lbls, dps)
    assert out.dtype == torch.int64
    assert np.array_equal(out.numpy().astype(np.uint8), g["consensus"])  # bit-exact


@pytest.mark.parametrize("name", golden_names("consensus"))
def test_cascade_model_reproduces_aten_sum_order(name):
    g = load_golden(name)
    n = g["labels"].shape[0]
    w = oc.softmax_weights([float(d) for d in g["dps"]]).numpy()
    weighted = g["labels"].reshape(n, -1).astype(np.float32) * w[:, None]
    votes = oc.cascade_sum_rows(weighted)
    ref_votes = (torch.from_numpy(g["labels"].astype(np.int64)) * torch.from_numpy(w).view(-1, 1, 1, 1)).sum(0)
    ref_votes = ref_votes.reshape(-1).numpy()
    # the vectorised-column main path is bit-exact; ATen's scalar tail columns (<32 per thread chunk)
    # use a 4-way interleave, so allow those few columns to differ by rounding only
    exact = votes == ref_votes
    assert exact.mean() > 0.9
    assert np.allclose(votes, ref_votes, rtol=1e-6, atol=0)
    assert np.array_equal((votes > 0.5)[exact], g["consensus"].reshape(-1).astype(bool)[exact])


@pytest.mark.parametrize("name", golden_names("loss"))
def test_loss_oracle_matches_reference_output(name):
    g = load_golden(name)
    res = ol.run_loss(
        torch.from_numpy(g["logits"]), torch.from_numpy(g["target"].astype(np.int64)),
        torch.from_numpy(g["emb"]), torch.from_numpy(g["idx"]),
        torch.from_numpy(g["fixed"]) if bool(g["use_fixed"]) else None,
        torch.from_numpy(g["cw"]), bool(g["risk"]), float(g["g"]), bool(g["need_dlogits"]),
    )
    # same ATen ops in the same order on the same torch build -> identical bits
    assert np.array_equal(res.ce_loss.numpy(), g["ce_loss"])
    assert np.array_equal(res.dp_loss.numpy(), g["dp_loss"])
    assert np.array_equal(res.d_bare_weight.numpy(), g["d_bare_weight"])
    assert res.d_embedding.is_sparse
    assert np.array_equal(res.d_logits_ce.numpy(), g["d_logits_ce"])
    if bool(g["need_dlogits"]):
        assert res.d_logits.dtype == torch.from_numpy(g["logits"]).dtype
        assert np.array_equal(res.d_logits.numpy(), g["d_logits"])
    else:
        assert res.d_logits is None


@pytest.mark.parametrize("name", golden_names("loss"))
def test_closed_form_gradients_match_reference_autograd(name):
    """the formulas the CUDA epilogue implements, checked against the
    reference's autograd output (1e-5 rel on the loss, 1e-4 rel on the gradients)."""
    g = load_golden(name)
    logits = torch.from_numpy(g["logits"]).double()
    target = torch.from_numpy(g["target"].astype(np.int64))
    idx = torch.from_numpy(g["idx"])
    n_dims = ol.spatial_dims(logits)
    per_inst = torch.nn.CrossEntropyLoss(reduction="none")(logits, target).mean(n_dims)
    fg = (logits.argmax(1) > 0).sum(n_dims)
    numel = math.prod(logits.shape[2:])
    bare = torch.from_numpy(g["emb"]).reshape(-1)[idx]
    fixed_b = torch.from_numpy(g["fixed"])[idx] if bool(g["use_fixed"]) else None
    loss, grad, weight = ol.closed_form(per_inst, fg, bare, fixed_b, numel, bool(g["risk"]), float(g["g"]))
    assert abs(loss.item() - float(g["dp_loss"])) <= 1e-5 * abs(float(g["dp_loss"])) + 1e-7
    scale = np.abs(g["d_bare_weight"]).max() + 1e-30
    assert np.abs(grad.numpy() - g["d_bare_weight"]).max() <= 1e-4 * scale + 1e-9 * float(g["g"])
    if bool(g["need_dlogits"]):
        p = torch.softmax(logits, 1)
        onehot = torch.nn.functional.one_hot(target, logits.shape[1]).movedim(-1, 1).double()
        shape = [-1] + [1] * (logits.dim() - 1)
        dz = float(g["g"]) * (weight / numel).reshape(shape) * (p - onehot)
        ref_dz = g["d_logits"].astype(np.float64)
        tol = 1e-4 if g["logits"].dtype == np.float32 else 2e-3  # fp16 grads carry fp16 rounding
        assert np.abs(dz.numpy() - ref_dz).max() <= tol * np.abs(ref_dz).max()


@pytest.mark.parametrize("name", golden_names("nonfinite"))
def test_nonfinite_oracle_matches_reference_output(name):
    """NaN / Inf logits (main_deep_staple.py:762-768 relies on them reaching the gradients so that the GradScaler
    skips the step): the restatement reproduces the reference's NaN / Inf pattern and every finite value."""
    g = load_golden(name)
    res = ol.run_loss(
        torch.from_numpy(g["logits"]), torch.from_numpy(g["target"].astype(np.int64)),
        torch.from_numpy(g["emb"]), torch.from_numpy(g["idx"]),
        torch.from_numpy(g["fixed"]) if bool(g["use_fixed"]) else None,
        torch.from_numpy(g["cw"]), bool(g["risk"]), float(g["g"]), bool(g["need_dlogits"]),
    )
    eq = lambda a, b: np.array_equal(a, b, equal_nan=True)
    assert eq(res.ce_loss.numpy(), g["ce_loss"]) and eq(res.dp_loss.numpy(), g["dp_loss"])
    assert eq(res.d_bare_weight.numpy(), g["d_bare_weight"])
    assert eq(res.per_inst_ce.numpy(), g["per_inst_ce"]) and eq(res.fg_pred.numpy(), g["fg_pred"])
    assert eq(res.d_logits_ce.numpy(), g["d_logits_ce"])
    if bool(g["need_dlogits"]):
        assert eq(res.d_logits.numpy(), g["d_logits"])
    # the closed form of SURVEY.md Appendix A (what the CUDA epilogue evaluates) propagates NaN / Inf identically
    bare = torch.from_numpy(g["emb"]).reshape(-1)[torch.from_numpy(g["idx"])]
    fixed_b = torch.from_numpy(g["fixed"])[torch.from_numpy(g["idx"])] if bool(g["use_fixed"]) else None
    numel = math.prod(g["logits"].shape[2:])
    loss, grad, _ = ol.closed_form(torch.from_numpy(g["per_inst_ce"]), torch.from_numpy(g["fg_pred"]), bare, fixed_b, numel,
                                   bool(g["risk"]), float(g["g"]))
    assert np.isnan(loss.item()) == np.isnan(g["dp_loss"]) and np.isinf(loss.item()) == np.isinf(g["dp_loss"])
    if g["logits"].shape[0] > 1:  # B == 1: the closed form is 0 x c_b, see the kernel epilogue
        assert np.array_equal(np.isnan(grad.numpy()), np.isnan(g["d_bare_weight"]))
        assert np.array_equal(np.isneginf(grad.numpy()), np.isneginf(g["d_bare_weight"]))


@pytest.mark.parametrize("name", golden_names("prepass"))
def test_prepass_oracle_matches_reference_output(name):
    g = load_golden(name)
    cw, fw, counts, gt = ol.prepass_constants([torch.from_numpy(l.astype(np.int64)) for l in g["labels"]], int(g["num_classes"]))
    assert np.array_equal(cw.numpy(), g["class_weights"])
    assert np.array_equal(fw.numpy(), g["fixed_weighting"])
    assert np.array_equal(counts.numpy(), g["counts"])
    assert np.array_equal(gt.numpy(), g["gt_num"])


@pytest.mark.parametrize("name", golden_names("dice"))
def test_dice_oracle_matches_reference_output(name):
    g = load_golden(name)
    d = om.dice_from_label_maps(torch.from_numpy(g["pred"]), torch.from_numpy(g["target"]), int(g["num_classes"]),
                                bool(g["nan_for_unlabeled_target"]))
    assert np.array_equal(d.numpy(), g["dice"], equal_nan=True)


@pytest.mark.skipif(not ref.available(), reason="reference tree not mounted (GPU box)")
def test_drift_guard_reference_lines_still_produce_the_golden_vectors():
    fn = ref.reference_calc_dp_consensus()
    g = load_golden("consensus__bernoulli_n30")
    out = fn([torch.from_numpy(l.astype(np.int64)) for l in g["labels"]],
             [torch.tensor([d], dtype=torch.float32) for d in g["dps"]])
    assert np.array_equal(out.numpy().astype(np.uint8), g["consensus"])
    g = load_golden("loss__b4_c2_3d")
    res = ref.reference_loss(torch.from_numpy(g["logits"]), torch.from_numpy(g["target"].astype(np.int64)),
                             torch.from_numpy(g["emb"]), torch.from_numpy(g["idx"]), torch.from_numpy(g["fixed"]),
                             torch.from_numpy(g["cw"]), True, True, 1.0, True)
    assert np.array_equal(res["dp_loss"].numpy(), g["dp_loss"])
    assert np.array_equal(res["d_logits"].numpy(), g["d_logits"])


def test_snapshot_grouping_and_driver():
    """consensus.ipynb:353-418 restated: ids 'XXXX:mYYYY' group by fixed id; vote per fixed id."""
    torch.manual_seed(0)
    shape = (4, 6, 8)
    ids = ["100l:m200l", "100l:m201r", "101r:m200l", "100l:m202l", "101r:m203r"]
    mod = torch.stack([(torch.rand(shape) < 0.4).long().to_sparse() for _ in ids])
    lab = torch.stack([(torch.rand(shape) < 0.4).long().to_sparse() for _ in ids])
    pred = torch.stack([(torch.rand(1, *shape) < 0.4).long().to_sparse() for _ in ids])
    snap = {"d_ids": tuple(ids), "data_parameters": torch.randn(len(ids), 1), "labels": lab,
            "modified_labels": mod, "train_predictions": pred}
    out = oc.consensus_from_snapshot(snap)
    assert list(out.keys()) == ["100l", "101r"]
    assert set(out["100l"].keys()) >= {"expert_label", "prediction", "200l", "201r", "202l", "dp_consensus",
                                       "dp_consensus_oracle_dice"}
    direct = oc.calc_dp_consensus([mod[0], mod[1], mod[3]], [snap["data_parameters"][i] for i in (0, 1, 3)])
    assert torch.equal(out["100l"]["dp_consensus"].to_dense(), direct)
    assert out["100l"]["dp_consensus_oracle_dice"].shape == (1, 2)
