"""Drop-in for the OOD half of the reference's ``lib/utils/metric.py`` (lines 66-180).

Same names, arguments and return conventions; the computation runs on the GPU
(hand-written sm_100a kernels behind libmss_b200.so) and is bit-identical to the
reference's numpy + scikit-learn result:

    eval_ood_measure(conf, seg_label, train_id_in=0, train_id_out=1) -> (auroc, aupr, fpr) | None
    get_and_print_results(out_score, in_score)                         -> (auroc, aupr, fpr)
    get_measures(_pos, _neg, recall_level=0.95)                        -> (auroc, aupr, fpr)
    fpr_and_fdr_at_recall(y_true, y_score, recall_level=0.95, pos_label=None) -> fpr

``conf`` / ``seg_label`` may be numpy arrays (what ``test_deeplab.py:98-101`` passes), CPU
tensors, or -- avoiding the device->host round trip of ``test_deeplab.py:94-95`` altogether --
CUDA tensors.  Scores must be float32 (or a narrower float type): thresholds are distinct
float32 values exactly as in the reference, whose testers only ever pass float32.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib as L

__all__ = ["eval_ood_measure", "get_and_print_results", "get_measures", "fpr_and_fdr_at_recall",
           "metrics_from_sorted_streams", "PairBuffer"]


def _device(*xs) -> torch.device:
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    if not torch.cuda.is_available():
        raise L.MssError("no CUDA device: multishiftseg_b200 computes this path on the GPU only")
    return torch.device("cuda", torch.cuda.current_device())


def _as_scores(x, device) -> torch.Tensor:
    t = torch.as_tensor(x)
    if t.dtype == torch.float64:
        raise TypeError("scores must be float32 (float64 has thresholds that float32 cannot represent); "
                        "the reference's testers pass float32 score maps")
    if not t.dtype.is_floating_point:
        raise TypeError(f"scores must be floating point, got {t.dtype}")
    t = t.to(device=device, non_blocking=True)
    if t.dtype != torch.float32:
        t = t.float()                       # exact widening of fp16 / bf16
    return t.reshape(-1).contiguous()


def _as_labels(x, device) -> torch.Tensor:
    t = torch.as_tensor(x).to(device=device, non_blocking=True)
    if t.dtype not in (torch.uint8, torch.int32, torch.int64):
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
        elif t.dtype.is_floating_point:
            raise TypeError("labels must be an integer tensor")
        else:
            t = t.to(torch.int64)
    return t.reshape(-1).contiguous()


class PairBuffer:
    """Device buffer of the order-preserving keys of valid pixels -- the on-device replacement of the testers'
    ``anomaly_scores`` / ``ood_gts`` host lists (test_deeplab.py:84-101, test_m2f.py:125-144).

    Two key-only streams share ``keys``: in-distribution keys fill ``keys[0:n_neg]`` upwards, OOD keys fill
    ``keys[capacity - n_pos:]`` downwards; the 0/1 label of a pixel is the stream its key is in (no label bytes)."""

    def __init__(self, capacity: int, device):
        self.device = torch.device(device)
        self.capacity = int(capacity)
        self.keys = torch.empty(max(self.capacity, 1), dtype=torch.int32, device=self.device)
        self.state = torch.zeros(L.EVAL_STATE_BYTES, dtype=torch.uint8, device=self.device)
        self.c = L.EvalBuffers(self.keys.data_ptr(), self.state.data_ptr(), self.capacity)

    @classmethod
    def from_tensors(cls, keys: torch.Tensor, state: torch.Tensor, capacity: int):
        """Wrap existing device memory (e.g. a peer-mapped receive buffer of the multi-GPU exchange)."""
        self = cls.__new__(cls)
        self.device = keys.device
        self.capacity = int(capacity)
        self.keys, self.state = keys, state
        self.c = L.EvalBuffers(keys.data_ptr(), state.data_ptr(), self.capacity)
        return self

    def reset(self):
        with torch.cuda.device(self.device):
            L.check(L.load().mss_eval_reset(C.byref(self.c), L.stream_ptr(self.device)), "mss_eval_reset")

    def append(self, scores: torch.Tensor, labels: torch.Tensor, id_in: int = 0, id_out: int = 1):
        assert scores.numel() == labels.numel(), "scores and labels differ in size"
        with torch.cuda.device(self.device):
            L.check(L.load().mss_eval_append(scores.data_ptr(), labels.data_ptr(), L.label_code(labels),
                                             scores.numel(), id_in, id_out, C.byref(self.c),
                                             L.stream_ptr(self.device)), "mss_eval_append")

    def streams(self, count: int, n_pos: int):
        """-> (negative keys, positive keys) views for a state (count, n_pos) read with ``read_state``."""
        return self.keys[: count - n_pos], self.keys[self.capacity - n_pos: self.capacity]

    def grow(self, capacity: int):
        """Enlarge (keeps the appended keys)."""
        if capacity <= self.capacity:
            return
        m, n_pos, _, _ = self.read_state()
        keys = torch.empty(capacity, dtype=torch.int32, device=self.device)
        neg, pos = self.streams(m, n_pos)
        keys[: m - n_pos].copy_(neg)
        keys[capacity - n_pos:].copy_(pos)
        self.keys, self.capacity = keys, int(capacity)
        self.c = L.EvalBuffers(self.keys.data_ptr(), self.state.data_ptr(), self.capacity)

    def read_state(self) -> Tuple[int, int, int, int]:
        """(count, n_pos, nan_flag, inf_flag); synchronises the stream."""
        out = (C.c_int64 * 4)()
        with torch.cuda.device(self.device):
            L.check(L.load().mss_eval_state_host(C.byref(self.c), out, L.stream_ptr(self.device)),
                    "mss_eval_state_host")
        return int(out[0]), int(out[1]), int(out[2]), int(out[3])

    def sort(self, n_upper: int):
        """Sort both streams in place (one sequence of launches; stream sizes are read from the device state)."""
        lib = L.load()
        with torch.cuda.device(self.device):
            nbytes = lib.mss_sort_keys_workspace_bytes(n_upper)
            ws = L.workspace(nbytes, self.device)
            L.check(lib.mss_eval_sort(C.byref(self.c), n_upper, ws.data_ptr(), nbytes, L.stream_ptr(self.device)),
                    "mss_eval_sort")
        del ws


def sort_keys(keys_a: torch.Tensor, n_a: int, keys_b: Optional[torch.Tensor] = None, n_b: int = 0):
    """Ascending in-place sort of one or two independent uint32 key arrays (int32 tensors holding the bits)."""
    dev = keys_a.device
    lib = L.load()
    with torch.cuda.device(dev):
        nbytes = lib.mss_sort_keys_workspace_bytes(n_a + n_b)
        ws = L.workspace(nbytes, dev)
        L.check(lib.mss_sort_keys(keys_a.data_ptr(), n_a, L.ptr(keys_b), n_b, ws.data_ptr(), nbytes, L.stream_ptr(dev)),
                "mss_sort_keys")
    del ws


def counts_from_sorted(neg_keys, n_neg: int, pos_keys, n_pos: int, pos_before: int = 0, neg_before: int = 0):
    """Two sorted key streams -> (tps[T], fps[T]) int64 device tensors (cumulative counts per distinct key)."""
    dev = neg_keys.device if neg_keys is not None else pos_keys.device
    lib = L.load()
    n = n_neg + n_pos
    with torch.cuda.device(dev):
        tps = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        fps = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        nbytes = lib.mss_counts_workspace_bytes(n)
        ws = L.workspace(nbytes, dev)
        T = C.c_int64(0)
        L.check(lib.mss_counts_from_sorted(L.ptr(neg_keys), n_neg, L.ptr(pos_keys), n_pos, pos_before, neg_before,
                                           tps.data_ptr(), fps.data_ptr(), C.byref(T), ws.data_ptr(), nbytes,
                                           L.stream_ptr(dev)), "mss_counts_from_sorted")
    return tps[:T.value], fps[:T.value]


def metrics_tail(tps: torch.Tensor, fps: torch.Tensor, recall_level: float = 0.95):
    """float64 tail over int64 cumulative counts -> ((auroc, ap, fpr), T_roc)."""
    dev = tps.device
    lib = L.load()
    T = tps.numel()
    assert fps.numel() == T and T >= 1
    tps = tps.contiguous()
    fps = fps.contiguous()
    with torch.cuda.device(dev):
        nbytes = lib.mss_tail_workspace_bytes(T)
        ws = L.workspace(nbytes, dev)
        out = (C.c_double * 3)()
        t_roc = C.c_int64(0)
        rc = L.check(lib.mss_metrics_tail(tps.data_ptr(), fps.data_ptr(), T, float(recall_level), ws.data_ptr(),
                                          nbytes, out, C.byref(t_roc), L.stream_ptr(dev)), "mss_metrics_tail")
    if rc == L.MSS_EMPTY_CLASS:
        return None, int(t_roc.value)
    return (np.float64(out[0]), np.float64(out[1]), np.float64(out[2])), int(t_roc.value)


def metrics_from_sorted_streams(neg_keys, n_neg: int, pos_keys, n_pos: int, recall_level: float = 0.95):
    tps, fps = counts_from_sorted(neg_keys, n_neg, pos_keys, n_pos)
    res, _ = metrics_tail(tps, fps, recall_level)
    return res


# up to this many keys the C composite runs (every stage sized for the upper bound, two host round trips in all);
# beyond it the stages run one by one so that the float64 tail is sized by the actual number of thresholds
_FUSED_MAX = 1 << 28


def _finish(buf: PairBuffer, recall_level: float = 0.95, check_empty: bool = True):
    m, n_pos, nan, inf = buf.read_state()
    if check_empty and (n_pos == 0 or n_pos == m):
        return None                                               # metric.py:176-180
    if nan:
        raise ValueError("Input contains NaN.")                   # sklearn assert_all_finite
    if inf:
        raise ValueError("Input contains infinity or a value too large for dtype('float32').")
    if m <= _FUSED_MAX and recall_level == 0.95:
        lib = L.load()
        with torch.cuda.device(buf.device):
            nbytes = lib.mss_ood_metrics_from_eval_workspace_bytes(m)
            ws = L.workspace(nbytes, buf.device)
            out = (C.c_double * 3)()
            rc = L.check(lib.mss_ood_metrics_from_eval(C.byref(buf.c), ws.data_ptr(), nbytes, out, None,
                                                       L.stream_ptr(buf.device)), "mss_ood_metrics_from_eval")
        if rc == L.MSS_EMPTY_CLASS:
            return None
        return np.float64(out[0]), np.float64(out[1]), np.float64(out[2])
    buf.sort(m)
    neg, pos = buf.streams(m, n_pos)
    return metrics_from_sorted_streams(neg, m - n_pos, pos, n_pos, recall_level)


def eval_ood_measure(conf, seg_label, train_id_in=0, train_id_out=1) -> Optional[Tuple[float, float, float]]:
    """metric.py:170-180.  Pixels whose label is neither ``train_id_in`` nor ``train_id_out`` (255 = ignore
    in every dataset of the reference, lib/dataset/anomaly.py) are dropped; ``None`` when a class is empty."""
    dev = _device(conf, seg_label)
    scores = _as_scores(conf, dev)
    labels = _as_labels(seg_label, dev)
    if scores.numel() != labels.numel():
        raise IndexError("conf and seg_label differ in size")     # numpy boolean-index error in the reference
    n = scores.numel()
    if n == 0:
        return None
    if n > _FUSED_MAX:
        buf = PairBuffer(n, dev)
        buf.reset()
        buf.append(scores, labels, int(train_id_in), int(train_id_out))
        return _finish(buf)
    # one C call: selection + key build + sort + counts + ROC compaction enqueued back to back, two host syncs in all
    lib = L.load()
    with torch.cuda.device(dev):
        nbytes = lib.mss_ood_metrics_workspace_bytes(n)
        ws = L.workspace(nbytes, dev)
        out = (C.c_double * 3)()
        rc = L.check(lib.mss_ood_metrics(scores.data_ptr(), labels.data_ptr(), L.label_code(labels), n, int(train_id_in),
                                         int(train_id_out), ws.data_ptr(), nbytes, out, None, L.stream_ptr(dev)),
                     "mss_ood_metrics")
    if rc == L.MSS_EMPTY_CLASS:
        return None
    return np.float64(out[0]), np.float64(out[1]), np.float64(out[2])


def get_measures(_pos, _neg, recall_level=0.95):
    """metric.py:130-153: positives (OOD scores) and negatives (ID scores) given separately."""
    dev = _device(_pos, _neg)
    pos = _as_scores(_pos, dev)
    neg = _as_scores(_neg, dev)
    buf = PairBuffer(pos.numel() + neg.numel(), dev)
    if pos.numel():
        buf.append(pos, torch.ones(pos.numel(), dtype=torch.uint8, device=dev))
    if neg.numel():
        buf.append(neg, torch.zeros(neg.numel(), dtype=torch.uint8, device=dev))
    res = _finish(buf, recall_level)
    if res is None:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    return res


def get_and_print_results(out_score, in_score):
    """metric.py:156-168 (the np.mean over one-element lists is the identity)."""
    return get_measures(out_score, in_score)


def fpr_and_fdr_at_recall(y_true, y_score, recall_level=0.95, pos_label=None):
    """metric.py:87-127: FPR at the threshold whose recall is closest to ``recall_level``."""
    dev = _device(y_true, y_score)
    yt = torch.as_tensor(y_true).to(dev).reshape(-1)
    classes = torch.unique(yt).tolist()
    if pos_label is None and classes not in ([0, 1], [-1, 1], [0], [-1], [1]):
        raise ValueError("Data is not binary and pos_label is not specified")
    if pos_label is None:
        pos_label = 1
    labels = (yt == pos_label).to(torch.uint8)
    scores = _as_scores(y_score, dev)
    buf = PairBuffer(scores.numel(), dev)
    buf.append(scores, labels)
    res = _finish(buf, recall_level)
    if res is None:
        raise ValueError("fpr_and_fdr_at_recall needs both classes present")
    return res[2]
