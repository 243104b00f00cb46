"""ctypes binding of libmss_b200.so (the C ABI in include/mss_b200.h).

There is NO fallback: if the library is missing, or the tensors are not on a CUDA
device, the calls raise.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmss_b200.so")

MSS_OK = 0
MSS_EMPTY_CLASS = 1
MSS_ERR_INVALID_ARG = -1
MSS_ERR_CUDA = -2
MSS_ERR_NAN = -3
MSS_ERR_INF = -4
MSS_ERR_WORKSPACE = -5
MSS_ERR_UNSUPPORTED = -6

LABEL_U8, LABEL_I32, LABEL_I64 = 1, 4, 8
SCORE_ENERGY, SCORE_MAXLOGIT, SCORE_MSP, SCORE_ENTROPY = 1, 2, 4, 8
SCORE_BITS = {"energy": SCORE_ENERGY, "maxlogit": SCORE_MAXLOGIT, "msp": SCORE_MSP, "entropy": SCORE_ENTROPY}
M2F_FORCE_GENERIC = 1
M2F_FORCE_FFMA = 2
M2F_FORCE_MMASYNC = 4
M2F_FORCE_TC5_PIXEL = 8
EVAL_STATE_BYTES = 64


class MssError(RuntimeError):
    pass


class EvalBuffers(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("state", C.c_void_p), ("capacity", C.c_int64)]


_p, _i, _i64, _u, _sz, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint, C.c_size_t, C.c_double
_EV = C.POINTER(EvalBuffers)

# name -> (restype, argtypes); must list every symbol include/mss_b200.h declares (tests check this)
SIGNATURES = {
    "mss_abi_version": (_i, []),
    "mss_last_error": (C.c_char_p, []),
    "mss_launch_count": (_i64, []),
    "mss_eval_reset": (_i, [_EV, _p]),
    "mss_eval_append": (_i, [_p, _p, _i, _i64, _i64, _i64, _EV, _p]),
    "mss_eval_state_host": (_i, [_EV, _p, _p]),
    "mss_deeplab_score": (_i, [_p, _i64, _i, _i64, _u, _p, _p, _p, _p, _p, _i, _i64, _i64, _u, _EV, _p]),
    "mss_upsample_bilinear": (_i, [_p, _i64, _i, _i, _p, _i, _i, _i, _p]),
    "mss_deeplab_anomaly_score": (_i, [_p, _i64, _i, _i, _i, _p, _p, _i, _i, _p]),
    "mss_deeplab_energy_backward": (_i, [_p, _p, _i64, _i, _i64, _p, _p]),
    "mss_upsample_bilinear_backward": (_i, [_p, _i64, _i, _i, _p, _i, _i, _i, _p]),
    "mss_deeplab_anomaly_score_backward": (_i, [_p, _p, _i64, _i, _i, _i, _i, _i, _p, _p]),
    "mss_deeplab_head_workspace_bytes": (_sz, [_i]),
    "mss_deeplab_head": (_i, [_p, _i64, _i, _i64, _p, _p, _i, _p, _p, _p, _p, _sz, _p]),
    "mss_m2f_workspace_bytes": (_sz, [_i64, _i, _i]),
    "mss_m2f_semantic_inference": (_i, [_p, _p, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i64, _p, _p, _p, _p,
                                        _p, _i64, _p, _sz, _u, _p]),
    "mss_m2f_anomaly_backward_workspace_bytes": (_sz, [_i64, _i, _i, _i, _i]),
    "mss_m2f_anomaly_backward": (_i, [_p, _p, _p, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "mss_m2f_mask_logits_workspace_bytes": (_sz, [_i64, _i]),
    "mss_m2f_mask_logits": (_i, [_p, _p, _i64, _i, _i, _i64, _p, _p, _sz, _p]),
    "mss_ood_metrics_workspace_bytes": (_sz, [_i64]),
    "mss_ood_metrics": (_i, [_p, _p, _i, _i64, _i64, _i64, _p, _sz, _p, _p, _p]),
    "mss_ood_metrics_from_eval_workspace_bytes": (_sz, [_i64]),
    "mss_ood_metrics_from_eval": (_i, [_EV, _p, _sz, _p, _p, _p]),
    "mss_ood_metrics_dist": (_i, [_EV, _p, _i, _i, _p, _p, _p]),
    "mss_sort_keys_workspace_bytes": (_sz, [_i64]),
    "mss_sort_keys": (_i, [_p, _i64, _p, _i64, _p, _sz, _p]),
    "mss_eval_sort": (_i, [_EV, _i64, _p, _sz, _p]),
    "mss_keys_histogram": (_i, [_p, _i64, _i, _p, _p]),
    "mss_keys_histogram_sampled": (_i, [_p, _i64, _i, _i, _p, _p]),
    "mss_keys_histogram_refine": (_i, [_p, _i64, _u, _i, _p, _p]),
    "mss_partition_workspace_bytes": (_sz, [_i64, _i]),
    "mss_partition_keys": (_i, [_p, _i64, _p, _i, _p, _p, _p, _sz, _p]),
    "mss_partition_count": (_i, [_p, _i64, _p, _i, _p, _p, _sz, _p]),
    "mss_partition_scatter_keys": (_i, [_p, _i64, _p, _i, _p, _p, _p, _sz, _p]),
    "mss_eval_partition_workspace_bytes": (_sz, [_i64, _i64, _i]),
    "mss_eval_partition_count": (_i, [_EV, _i64, _i64, _p, _i, _p, _p, _sz, _p]),
    "mss_eval_partition_scatter": (_i, [_EV, _i64, _i64, _p, _i, _p, _p, _p, _p, _sz, _p]),
    "mss_eval_exchange_append": (_i, [_EV, _i64, _i64, _p, _i, _p, _p, _i64, _p, _sz, _p]),
    "mss_eval_exchange_stream_workspace_bytes": (_sz, [_i64, _i]),
    "mss_eval_exchange_stream": (_i, [_EV, _p, _i, _p, _p, _i64, _p, _p, _sz, _p]),
    "mss_eval_exchange_stage": (_i, [_EV, _p, _i, _p, _p, _i64, _p, _i64, _p, _p, _p]),
    "mss_memcpy_async": (_i, [_p, _p, _sz, _p]),
    "mss_counts_workspace_bytes": (_sz, [_i64]),
    "mss_counts_from_sorted": (_i, [_p, _i64, _p, _i64, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    "mss_tail_workspace_bytes": (_sz, [_i64]),
    "mss_metrics_tail": (_i, [_p, _p, _i64, _d, _p, _sz, _p, _p, _p]),
    "mss_pairwise_leaf_bounds": (_i, [_i64, _i64, _p, _p, _p]),
    "mss_pairwise_sum_host": (_i, [_p, _i64, _p]),
    "mss_confusion_hist": (_i, [_p, _i, _p, _i, _i64, _i, _p, _p, _p]),
    "mss_confusion_from_logits": (_i, [_p, _i64, _i, _i64, _p, _i, _i, _p, _p, _p]),
    "mss_confusion_result": (_i, [_p, _p, _i, _p, _p, _p]),
    "mss_confusion_scores": (_i, [_p, _i, _d, _d, _i, _p, _p, _p]),
    "mss_deeplab_score_host_scratch_bytes": (_sz, [_i64, _i, _i64, _u]),
    "mss_deeplab_score_host": (_i, [_p, _i64, _i, _i64, _u, _p, _p, _p, _p, _p, _sz, _p]),
}

_lib = None
_lock = threading.Lock()


def load():
    """dlopen libmss_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise MssError(
                    f"{LIB_PATH} is missing: build it with `python -m multishiftseg_b200.build` "
                    "(nvcc, sm_100a).  There is no CPU / PyTorch fallback for this path.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def last_error() -> str:
    return load().mss_last_error().decode(errors="replace")


def check(rc: int, what: str) -> int:
    """Map C return codes to the reference's behaviour (SURVEY 8b error convention)."""
    if rc >= 0:
        return rc
    msg = last_error()
    if rc in (MSS_ERR_NAN, MSS_ERR_INF):
        raise ValueError(msg)                      # sklearn assert_all_finite (K10 / K11)
    if rc == MSS_ERR_INVALID_ARG:
        raise ValueError(f"{what}: {msg}")
    raise MssError(f"{what} failed ({rc}): {msg}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise MssError(f"{name} must be a CUDA tensor: this path runs on the GPU only (no CPU fallback)")
    return t


def forbid_grad(what: str, *tensors):
    """Entry points without a backward kernel refuse inputs that autograd is tracking: returning a tensor with no
    grad_fn would let a training loop (train_m2f.py:443 -> criterion -> loss.backward()) run with silently missing
    gradients.  Wrap the call in torch.no_grad() for inference, or detach the inputs."""
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise MssError(f"{what} has no backward kernel: its inputs require grad and autograd is enabled. "
                       "Call it under torch.no_grad() (evaluation), or detach() the inputs; the differentiable "
                       "entry points are deeplab.energy_func / Upsample / anomaly_score.")


def wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def ptr(t):
    return 0 if t is None else t.data_ptr()


def label_code(t: torch.Tensor) -> int:
    if t.dtype == torch.uint8:
        return LABEL_U8
    if t.dtype == torch.int32:
        return LABEL_I32
    if t.dtype == torch.int64:
        return LABEL_I64
    raise TypeError(f"labels must be uint8, int32 or int64, got {t.dtype}")


def workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def launch_count() -> int:
    return int(load().mss_launch_count())
