"""CPU oracle for the DeepSTAPLE voxel-wise label-weighting hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deep_staple_b200/`` may import this
package: the product path is the sm_100a CUDA library behind ``include/deep_staple_b200.h``
and it fails loudly when that library or a B200 is missing.  The only callers
allowed are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` (where the oracle is the thing being
*timed as the baseline*, never the thing shipped).

What it restates (reference = multimodallearning/deep_staple, pure Python/PyTorch):

* ``oracle.consensus``  -- ``calc_dp_consensus`` (deep_staple/consensus/consensus.ipynb:100-109)
  and the consensus driver loop (consensus.ipynb:353-418, minus the SimpleITK STAPLE baseline).
* ``oracle.dp_loss``    -- the inline data-parameter loss of ``train_DL``
  (main_deep_staple.py:716, 738-759, 762) and its closed-form gradients.
* ``oracle.metrics``    -- hard Dice (deep_staple/metrics.py:68-111) and the per-fold
  pre-pass constants (main_deep_staple.py:646-653).
* ``oracle.reference_lines`` -- executes the reference's OWN source lines when
  ``/root/reference`` is mounted (golden-vector generation + drift guard only).

Arithmetic is third-party PyTorch ATen in the reference (pinned torch 1.11.0,
poetry.lock:1739-1740); the oracle therefore uses torch CPU ops in the same order,
so fp32 results are the reference's results on this container's torch.

Parity pinning: the reference has no tests, fixtures or golden vectors for this path
(SURVEY.md section 4, 8c).  The oracle is pinned instead against outputs of the reference's
own code executed in the build container (``tests/golden/make_golden.py`` ->
``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).
"""
