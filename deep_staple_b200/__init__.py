"""deep_staple_b200 -- B200-native (sm_100a) implementation of DeepSTAPLE's voxel-wise label-weighting
hot path: the data-parameter-weighted segmentation loss (main_deep_staple.py:716-775 of
multimodallearning/deep_staple) and the consensus weighted vote (consensus.ipynb:100-109), behind the
reference's own Python call sites.  Hand-written CUDA kernels in a C-ABI shared library
(include/deep_staple_b200.h), bound with ctypes.  No CPU fallback.
"""
from ._binding import DeepStapleError, launch_count, lib_path, load  # noqa: F401
from .consensus import (  # noqa: F401
    calc_dp_consensus,
    calc_staple_consensus,
    consensus_batched,
    consensus_from_snapshot,
    consensus_multiclass,
    consensus_sparse,
    group_snapshot,
    staple_batched,
)
from .host_io import LabelUploader, narrow_labels_host  # noqa: F401
from .loss import (  # noqa: F401
    DPLossOutput,
    backward_logits,
    dp_weighted_loss,
    label_statistics,
    loss_statistics,
    weighted_ce,
)
from .metrics import dice_from_label_maps, dice_from_stats, prepass_constants  # noqa: F401
from .optim import LossSection, dp_optimizer_step  # noqa: F401
from .snapshot import build_snapshot, load_snapshot, save_snapshot  # noqa: F401
from . import distributed  # noqa: F401,E402

__version__ = "0.1.0"
