#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the REFERENCE'S OWN code.

Run in the build container (``/root/reference`` mounted read-only):

    python tests/golden/make_golden.py

Every fixture holds seeded synthetic inputs and the outputs produced by

* ``calc_dp_consensus`` exec-ed from deep_staple/consensus/consensus.ipynb:100-109,
* main_deep_staple.py:716, :719, :738-759, :762 exec-ed verbatim (``oracle.reference_lines``),
* main_deep_staple.py:646-653 (pre-pass constants),
* ``deep_staple/metrics.py`` imported from the reference tree (hard Dice),

on torch CPU.  The fixtures travel to the GPU box (the reference tree does not); tests compare
both the oracle restatement and the CUDA path against them.  ``manifest.json`` records the torch
version and the reference file hashes the vectors were produced from.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import reference_lines as ref  # noqa: E402


def rng(seed):
    return np.random.default_rng(seed)


def ellipsoid_labels(n, shape, seed, radius=(2.0, 5.0), jitter=1.5):
    """cfg1(ii)-style labels: one ellipsoid per candidate, centre jittered around the middle."""
    r = rng(seed)
    zz, yy, xx = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    out = np.zeros((n, *shape), np.uint8)
    for i in range(n):
        c = np.array(shape, np.float32) / 2 + r.normal(0, jitter, 3)
        rad = r.uniform(*radius, 3)
        out[i] = (((zz - c[0]) / rad[0]) ** 2 + ((yy - c[1]) / rad[1]) ** 2 + ((xx - c[2]) / rad[2]) ** 2) <= 1.0
    return out


def consensus_cases():
    r = rng(10)
    cases = {}
    cases["bernoulli_n10"] = dict(labels=(r.random((10, 20, 18, 16)) < 0.3).astype(np.uint8),
                                  dps=(2 * r.standard_normal(10)).astype(np.float32), sparse=False)
    cases["uniform_dp_n10_ties"] = dict(labels=(r.random((10, 12, 10, 16)) < 0.5).astype(np.uint8),
                                        dps=np.zeros(10, np.float32), sparse=False)
    cases["ellipsoid_n30"] = dict(labels=ellipsoid_labels(30, (16, 16, 16), 11),
                                  dps=(2 * r.standard_normal(30)).astype(np.float32), sparse=True)
    cases["bernoulli_n30"] = dict(labels=(r.random((30, 8, 16, 16)) < 0.45).astype(np.uint8),
                                  dps=(0.5 * r.standard_normal(30)).astype(np.float32), sparse=False)
    cases["sparse_n3"] = dict(labels=(r.random((3, 9, 8, 16)) < 0.2).astype(np.uint8),
                              dps=(r.standard_normal(3)).astype(np.float32), sparse=True)
    cases["single_n1"] = dict(labels=(r.random((1, 4, 8, 8)) < 0.5).astype(np.uint8),
                              dps=np.array([0.3], np.float32), sparse=False)
    cases["odd_shape_n17"] = dict(labels=(r.random((17, 7, 5, 3)) < 0.5).astype(np.uint8),
                                  dps=(3 * r.standard_normal(17)).astype(np.float32), sparse=False)
    cases["n40"] = dict(labels=(r.random((40, 4, 8, 16)) < 0.5).astype(np.uint8),
                        dps=(0.2 * r.standard_normal(40)).astype(np.float32), sparse=False)
    return cases


def make_consensus(out):
    fn = ref.reference_calc_dp_consensus()
    for name, case in consensus_cases().items():
        lbls = [torch.from_numpy(l.astype(np.int64)) for l in case["labels"]]
        if case["sparse"]:
            lbls = [l.to_sparse() for l in lbls]
        # the notebook passes a list of [1]-shaped fp32 tensors (consensus.ipynb:372, 389);
        # one case uses plain floats to cover that input form too
        if name == "single_n1":
            dps = [float(d) for d in case["dps"]]
        else:
            dps = [torch.tensor([d], dtype=torch.float32) for d in case["dps"]]
        res = fn(lbls, dps)
        assert res.dtype == torch.int64
        out[f"consensus__{name}"] = dict(
            labels=case["labels"], dps=case["dps"], sparse=np.array(case["sparse"]),
            consensus=res.numpy().astype(np.uint8),
        )


def loss_cases():
    cases = {}

    def mk(seed, b, c, spatial, m, dtype=np.float32, p_fg=0.2):
        r = rng(seed)
        logits = (2 * r.standard_normal((b, c, *spatial))).astype(np.float32)
        probs = np.array([1 - p_fg] + [p_fg / (c - 1)] * (c - 1))
        target = r.choice(c, size=(b, *spatial), p=probs).astype(np.uint8)
        emb = r.standard_normal((m, 1)).astype(np.float32)
        idx = r.permutation(m)[:b].astype(np.int64)
        gt_num = r.integers(0, 50000, m).astype(np.float32)
        fixed = (np.log(gt_num + np.float32(np.e)) + np.float32(np.e)).astype(np.float32)
        counts = np.bincount(target.reshape(-1), minlength=c).astype(np.float32)
        cw = 1.0 / np.power(counts, 0.35)
        cw = (cw / cw.mean()).astype(np.float32)
        return dict(logits=logits.astype(dtype), target=target, emb=emb, idx=idx, fixed=fixed, cw=cw)

    cases["b4_c2_3d"] = dict(mk(20, 4, 2, (12, 10, 8), 11), risk=True, use_fixed=True, g=1.0, dlogits=True)
    cases["b3_c3_3d"] = dict(mk(21, 3, 3, (6, 10, 8), 9), risk=True, use_fixed=True, g=1.0, dlogits=True)
    cases["b6_c2_2d"] = dict(mk(22, 6, 2, (16, 12), 40), risk=True, use_fixed=True, g=1.0, dlogits=True)
    cases["b4_norisk_nofixed"] = dict(mk(23, 4, 2, (6, 6, 8), 7), risk=False, use_fixed=False, g=1.0, dlogits=True)
    cases["b4_risk_nofixed"] = dict(mk(24, 4, 2, (6, 6, 8), 7), risk=True, use_fixed=False, g=1.0, dlogits=True)
    cases["b1_squeeze"] = dict(mk(25, 1, 2, (8, 8, 8), 5), risk=True, use_fixed=True, g=1.0, dlogits=True)
    cases["b5_scale1024_ool"] = dict(mk(26, 5, 2, (5, 7, 9), 13), risk=True, use_fixed=True, g=1024.0, dlogits=False)
    cases["b4_fp16_scale65536"] = dict(mk(27, 4, 2, (8, 8, 8), 10, np.float16), risk=True, use_fixed=True,
                                       g=65536.0, dlogits=True)
    cases["b2_c4_3d"] = dict(mk(28, 2, 4, (4, 6, 8), 6), risk=True, use_fixed=True, g=1.0, dlogits=True)
    return cases


def make_loss(out):
    for name, c in loss_cases().items():
        logits = torch.from_numpy(c["logits"])
        target = torch.from_numpy(c["target"].astype(np.int64))
        res = ref.reference_loss(
            logits, target, torch.from_numpy(c["emb"]), torch.from_numpy(c["idx"]),
            torch.from_numpy(c["fixed"]) if c["use_fixed"] else None,
            torch.from_numpy(c["cw"]), c["risk"], c["use_fixed"], c["g"], c["dlogits"],
        )
        assert res["d_embedding_is_sparse"]
        rec = dict(
            logits=c["logits"], target=c["target"], emb=c["emb"], idx=c["idx"], fixed=c["fixed"], cw=c["cw"],
            risk=np.array(c["risk"]), use_fixed=np.array(c["use_fixed"]), g=np.array(c["g"], np.float64),
            need_dlogits=np.array(c["dlogits"]),
            ce_loss=res["ce_loss"].numpy(), dp_loss=res["dp_loss"].numpy(),
            d_bare_weight=res["d_bare_weight"].numpy(),
            d_logits_ce=res["d_logits_ce"].numpy(),
        )
        if c["dlogits"]:
            rec["d_logits"] = res["d_logits"].numpy()
        out[f"loss__{name}"] = rec


def nonfinite_cases():
    """Non-finite logits: the reference's AMP flow relies on NaN / Inf reaching the loss and the gradients so that
    ``scaler_dp.step(optimizer_dp)`` skips the step (main_deep_staple.py:762-768).  Each case is a small seeded
    batch with a few voxels overwritten; ``where`` lists (b, c, flat voxel, value)."""
    cases = {}
    inf, nan = np.float32(np.inf), np.float32(np.nan)
    neg_nan = np.array([0xFFC00000], np.uint32).view(np.float32)[0]  # what x86 inf - inf produces

    def mk(seed, b, c, spatial, dtype, edits, dlogits=True, g=1024.0, use_fixed=True, risk=True):
        r = rng(seed)
        logits = (2 * r.standard_normal((b, c, *spatial))).astype(np.float32).astype(dtype)
        target = r.choice(c, size=(b, *spatial), p=[0.8] + [0.2 / (c - 1)] * (c - 1)).astype(np.uint8)
        flat = logits.reshape(b, c, -1)
        tflat = target.reshape(b, -1)
        for (bi, ci, vi, val, tgt) in edits:
            flat[bi, ci, vi] = val
            if tgt is not None:
                tflat[bi, vi] = tgt
        m = b + 3
        emb = r.standard_normal((m, 1)).astype(np.float32)
        idx = r.permutation(m)[:b].astype(np.int64)
        fixed = (np.log(r.integers(0, 50000, m).astype(np.float32) + np.float32(np.e)) + np.float32(np.e)).astype(np.float32)
        cw = (r.random(c) + 0.5).astype(np.float32)
        return dict(logits=logits, target=target, emb=emb, idx=idx, fixed=fixed, cw=cw, risk=risk, use_fixed=use_fixed,
                    g=g, dlogits=dlogits)

    f32, f16 = np.float32, np.float16
    # (instance, class, flat voxel, value, label forced at that voxel)
    cases["posinf_target"] = mk(40, 3, 2, (4, 8, 8), f32, [(1, 1, 37, inf, 1)])
    cases["posinf_other"] = mk(41, 3, 2, (4, 8, 8), f32, [(0, 0, 5, inf, 1)])
    cases["posinf_both"] = mk(42, 2, 2, (4, 8, 8), f32, [(1, 0, 9, inf, 0), (1, 1, 9, inf, 0)])
    cases["neginf_target"] = mk(43, 3, 2, (4, 8, 8), f32, [(2, 1, 100, -inf, 1)])
    cases["neginf_other_is_finite"] = mk(44, 3, 2, (4, 8, 8), f32, [(2, 0, 100, -inf, 1), (0, 1, 3, -inf, 0)])
    cases["neginf_both"] = mk(45, 3, 2, (4, 8, 8), f32, [(0, 0, 64, -inf, 1), (0, 1, 64, -inf, 1)])
    cases["nan_one_class"] = mk(46, 4, 2, (4, 8, 8), f32, [(3, 1, 255, nan, 0)])
    cases["negnan_ool"] = mk(47, 4, 2, (4, 8, 8), f32, [(0, 0, 0, neg_nan, 0)], dlogits=False)
    cases["posinf_ool"] = mk(48, 3, 2, (4, 8, 8), f32, [(1, 1, 17, inf, 0)], dlogits=False)
    cases["neginf_target_ool"] = mk(49, 3, 2, (4, 8, 8), f32, [(1, 0, 17, -inf, 0)], dlogits=False)
    cases["fp16_overflow"] = mk(50, 4, 2, (4, 8, 8), f16, [(2, 1, 77, 70000.0, 1), (2, 0, 78, -70000.0, 0)], g=65536.0)
    cases["fp16_overflow_ool"] = mk(51, 4, 2, (4, 8, 8), f16, [(1, 0, 7, 70000.0, 1)], dlogits=False, g=65536.0)
    cases["fp16_max_is_finite"] = mk(52, 2, 2, (4, 8, 8), f16, [(1, 0, 7, 65504.0, 1), (1, 1, 7, -65504.0, 1),
                                                             (0, 1, 8, 65504.0, 1)], g=65536.0)
    cases["c3_posinf"] = mk(53, 3, 3, (4, 8, 8), f32, [(1, 2, 11, inf, 2)])
    cases["c3_nan_argmax"] = mk(54, 3, 3, (4, 8, 8), f32, [(1, 1, 11, nan, 0), (2, 2, 5, inf, 0), (0, 0, 3, nan, 1)])
    cases["c3_neginf_target"] = mk(55, 3, 3, (4, 8, 8), f32, [(0, 1, 200, -inf, 1)])
    cases["c3_neginf_all"] = mk(56, 3, 3, (4, 8, 8), f32, [(0, 0, 200, -inf, 1), (0, 1, 200, -inf, 1), (0, 2, 200, -inf, 1)])
    cases["c3_neginf_other_is_finite"] = mk(57, 3, 3, (4, 8, 8), f32, [(0, 0, 20, -inf, 1), (0, 2, 20, -inf, 1)])
    cases["b1_nan"] = mk(58, 1, 2, (4, 8, 8), f32, [(0, 1, 1, nan, 1)])
    cases["b1_neginf_target"] = mk(59, 1, 2, (4, 8, 8), f32, [(0, 1, 1, -inf, 1)])
    cases["odd_shape_nan"] = mk(60, 2, 2, (3, 5, 7), f32, [(1, 0, 104, nan, 0)])
    cases["huge_finite_f32"] = mk(61, 2, 2, (4, 8, 8), f32, [(0, 1, 5, 3.0e5, 0), (1, 0, 6, -1.0e6, 1)])
    cases["norisk_nofixed_posinf"] = mk(62, 3, 2, (4, 8, 8), f32, [(1, 1, 37, inf, 1)], use_fixed=False, risk=False)
    return cases


def make_nonfinite(out):
    for name, c in nonfinite_cases().items():
        logits = torch.from_numpy(c["logits"])
        target = torch.from_numpy(c["target"].astype(np.int64))
        res = ref.reference_loss(
            logits, target, torch.from_numpy(c["emb"]), torch.from_numpy(c["idx"]),
            torch.from_numpy(c["fixed"]) if c["use_fixed"] else None,
            torch.from_numpy(c["cw"]), c["risk"], c["use_fixed"], c["g"], c["dlogits"],
        )
        z = logits.float()
        n_dims = tuple(range(-(z.dim() - 2), 0))  # main_deep_staple.py:513-516
        rec = dict(
            logits=c["logits"], target=c["target"], emb=c["emb"], idx=c["idx"], fixed=c["fixed"], cw=c["cw"],
            risk=np.array(c["risk"]), use_fixed=np.array(c["use_fixed"]), g=np.array(c["g"], np.float64),
            need_dlogits=np.array(c["dlogits"]),
            ce_loss=res["ce_loss"].numpy(), dp_loss=res["dp_loss"].numpy(),
            d_bare_weight=res["d_bare_weight"].numpy(), d_logits_ce=res["d_logits_ce"].numpy(),
            # main_deep_staple.py:738-739 and :751 on their own, for the per-instance outputs
            per_inst_ce=torch.nn.CrossEntropyLoss(reduction="none")(z, target).mean(n_dims).numpy(),
            fg_pred=(z.argmax(1) > 0).sum(dim=n_dims).numpy(),
        )
        if c["dlogits"]:
            rec["d_logits"] = res["d_logits"].numpy()
        out[f"nonfinite__{name}"] = rec


def make_prepass_and_dice(out):
    r = rng(30)
    labels = (r.random((5, 6, 8, 8)) < np.array([0.05, 0.3, 0.0, 0.6, 0.15]).reshape(5, 1, 1, 1)).astype(np.uint8)
    labels3 = r.choice(3, size=(4, 6, 8, 8), p=[0.9, 0.07, 0.03]).astype(np.uint8)
    for name, lab, c in (("binary", labels, 2), ("three_class", labels3, 3)):
        cw, fw, counts, gt = ref.reference_prepass([torch.from_numpy(l.astype(np.int64)) for l in lab], c)
        out[f"prepass__{name}"] = dict(labels=lab, num_classes=np.array(c), class_weights=cw.numpy(),
                                       fixed_weighting=fw.numpy(), counts=counts.numpy(), gt_num=gt.numpy())
    metrics = ref.reference_metrics_module()
    oh = torch.nn.functional.one_hot
    pred = r.choice(3, size=(4, 6, 8, 8), p=[0.85, 0.1, 0.05]).astype(np.uint8)
    pred[3] = 0
    tgt = labels3.copy()
    tgt[3] = 0  # instance with nothing labelled and nothing predicted in classes 1,2 -> NaN / 0
    for nan_flag in (True, False):
        d3 = metrics.dice3d(oh(torch.from_numpy(pred.astype(np.int64)), 3), oh(torch.from_numpy(tgt.astype(np.int64)), 3),
                            one_hot_torch_style=True, nan_for_unlabeled_target=nan_flag)
        out[f"dice__3d_nan{int(nan_flag)}"] = dict(pred=pred, target=tgt, num_classes=np.array(3),
                                                   nan_for_unlabeled_target=np.array(nan_flag), dice=d3.numpy())
    pred2 = (r.random((5, 16, 12)) < 0.3).astype(np.uint8)
    tgt2 = (r.random((5, 16, 12)) < 0.3).astype(np.uint8)
    d2 = metrics.dice2d(oh(torch.from_numpy(pred2.astype(np.int64)), 2), oh(torch.from_numpy(tgt2.astype(np.int64)), 2),
                        one_hot_torch_style=True)
    out["dice__2d"] = dict(pred=pred2, target=tgt2, num_classes=np.array(2),
                           nan_for_unlabeled_target=np.array(True), dice=d2.numpy())


def sha256(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def main():
    if not ref.available():
        raise SystemExit("the reference tree is not mounted; golden vectors can only be generated in the build container")
    torch.manual_seed(0)
    torch.set_num_threads(1)  # make the run independent of the host's thread count
    out = {}
    make_consensus(out)
    make_loss(out)
    make_nonfinite(out)
    make_prepass_and_dice(out)
    for name, rec in out.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    manifest = {
        "generator": "tests/golden/make_golden.py",
        "torch": torch.__version__,
        "numpy": np.__version__,
        "reference_files": {
            ref.MAIN: sha256(os.path.join(ref.REFERENCE_ROOT, ref.MAIN)),
            ref.NOTEBOOK: sha256(os.path.join(ref.REFERENCE_ROOT, ref.NOTEBOOK)),
            "deep_staple/metrics.py": sha256(os.path.join(ref.REFERENCE_ROOT, "deep_staple/metrics.py")),
        },
        "fixtures": sorted(out.keys()),
    }
    with open(os.path.join(HERE, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    print(f"wrote {len(out)} fixtures to {HERE}")


if __name__ == "__main__":
    main()
