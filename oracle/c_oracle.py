"""ctypes binding of the plain-C oracle (oracle/oracle_metrics.c).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_metrics.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle_metrics.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle_metrics.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.oracle_pairwise_sum.restype = ctypes.c_double
        L.oracle_pairwise_sum.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        L.oracle_ood_metrics.restype = ctypes.c_int
        L.oracle_ood_metrics.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                         ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_metrics_from_counts.restype = None
        L.oracle_metrics_from_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                 ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
        _lib = L
    return _lib


def pairwise_sum(a) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return float(lib().oracle_pairwise_sum(a.ctypes.data, a.size))


def _labels(seg_label):
    lab = np.ascontiguousarray(seg_label).ravel()
    if lab.dtype == np.uint8:
        return lab, 1
    if lab.dtype == np.int32:
        return lab, 4
    return np.ascontiguousarray(lab, dtype=np.int64), 8


def eval_ood_measure(conf, seg_label, train_id_in=0, train_id_out=1, return_counts=False):
    """C restatement of metric.py:170-180; None on an empty class, ValueError on NaN/Inf."""
    s = np.ascontiguousarray(conf, dtype=np.float32).ravel()
    lab, nbytes = _labels(seg_label)
    assert s.size == lab.size
    out = np.zeros(3, dtype=np.float64)
    counts = np.zeros(4, dtype=np.int64)
    rc = lib().oracle_ood_metrics(s.ctypes.data, lab.ctypes.data, nbytes, s.size,
                                  train_id_in, train_id_out, out.ctypes.data, counts.ctypes.data)
    if rc == 1:
        return None
    if rc == -1:
        raise ValueError("Input contains NaN.")
    if rc == -2:
        raise ValueError("Input contains infinity or a value too large for dtype('float32').")
    res = (float(out[0]), float(out[1]), float(out[2]))
    return (res, counts) if return_counts else res


def metrics_from_counts(tps, fps, recall_level=0.95):
    tps = np.ascontiguousarray(tps, dtype=np.int64)
    fps = np.ascontiguousarray(fps, dtype=np.int64)
    out = np.zeros(3, dtype=np.float64)
    aux = np.zeros(2, dtype=np.int64)
    lib().oracle_metrics_from_counts(tps.ctypes.data, fps.ctypes.data, tps.size, recall_level,
                                     out.ctypes.data, aux.ctypes.data)
    return float(out[0]), float(out[1]), float(out[2])
