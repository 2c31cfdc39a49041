"""ctypes binding of libdeep_staple_b200.so (C ABI declared in include/deep_staple_b200.h).

The library is the product: there is no eager / CPU fallback anywhere in this package.  Loading
fails loudly when the shared object is missing, and every compute entry point fails loudly on a
device that is not sm_100 class.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

import torch

# DS_LIB_PATH: kernel-tuning experiments load an alternative build of the SAME library (build.py --variant)
_LIB_PATH = os.environ.get("DS_LIB_PATH") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "_lib", "libdeep_staple_b200.so")

# ds_dtype (include/deep_staple_b200.h)
DS_F32, DS_F16, DS_BF16, DS_I64, DS_U8 = 0, 1, 2, 3, 4
DS_STATS, DS_STAT_NLL, DS_STAT_LABEL, DS_STAT_PRED, DS_STAT_TP = 4, 0, 1, 2, 3
DS_INST, DS_INST_CE, DS_INST_FG, DS_INST_WEIGHT, DS_INST_SIGMOID = 4, 0, 1, 2, 3
DS_MAX_CLASSES, DS_MAX_CANDIDATES, DS_MAX_BATCH, DS_STAPLE_MAX_RATERS = 8, 65535, 4096, 64

_LOGIT_DTYPES = {torch.float32: DS_F32, torch.float16: DS_F16, torch.bfloat16: DS_BF16}
_LABEL_DTYPES = {torch.int64: DS_I64, torch.uint8: DS_U8}

# every symbol include/deep_staple_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "ds_version": (c_int, []),
    "ds_last_error_string": (c_char_p, []),
    "ds_device_check": (c_int, []),
    "ds_launch_count": (c_uint64, []),
    "ds_consensus_vote": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_int,
                                  c_void_p, c_void_p]),
    "ds_consensus_vote_gather": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int64,
                                         c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ds_consensus_vote_multiclass": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_int64, c_int,
                                             c_void_p, c_int, c_void_p]),
    "ds_consensus_vote_sparse": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64,
                                         c_void_p, c_int, c_void_p, c_void_p]),
    "ds_staple_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ds_staple": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                          c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ds_dp_loss_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ds_dp_loss_fwd_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                   c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ds_dp_loss_fwd_bwd_indexed": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ds_dp_loss_exchange_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "ds_dp_loss_fwd_bwd_exchange": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64,
                                            c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_uint64,
                                            c_void_p, c_size_t, c_void_p]),
    "ds_dp_loss_finalize": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ds_weighted_ce_fwd_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                       c_void_p]),
    "ds_dp_loss_backward_logits": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p,
                                           c_int64, c_void_p, c_void_p, c_void_p]),
    "ds_rescale_inplace": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "ds_dp_loss_backward_reconcile": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                              c_void_p]),
    "ds_sparse_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                                    c_double, c_double, c_double, c_double, c_void_p]),
    "ds_sparse_adam_scaler_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_double, c_double, c_double, c_double, c_double,
                                           c_double, c_int, c_void_p]),
    "ds_label_stats": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                               c_size_t, c_void_p]),
    "ds_host_narrow_labels": (c_int, [c_void_p, c_void_p, c_int64, c_int]),
}

_lib = None


def lib_path() -> str:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """Load the shared library (once).  No fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -m deep_staple_b200.build` "
            "(deep_staple_b200 has no CPU or eager fallback)"
        )
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class DeepStapleError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = load().ds_last_error_string()
    msg = msg.decode() if msg else ""
    if rc in (-2, -3, -6, -1):
        raise ValueError(f"{what}: {msg} (ds_status {rc})")
    raise DeepStapleError(f"{what}: {msg} (code {rc})")


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    """Raw handle of the caller's current stream on `device` (no Stream object is built: ~0.3 us instead of ~3)."""
    return torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device())


def logit_dtype(t: torch.Tensor) -> int:
    try:
        return _LOGIT_DTYPES[t.dtype]
    except KeyError:
        raise ValueError(f"logits must be float32, float16 or bfloat16, got {t.dtype}") from None


def label_dtype(t: torch.Tensor) -> int:
    try:
        return _LABEL_DTYPES[t.dtype]
    except KeyError:
        raise ValueError(f"labels must be int64 or uint8, got {t.dtype}") from None


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: deep_staple_b200 runs on B200 (sm_100a) only and has no CPU fallback"
        )


def launch_count() -> int:
    return int(load().ds_launch_count())
