"""Drop-in for the mIoU half of the reference's ``lib/utils/metric.py`` (lines 10-64), SURVEY 8f row 3.

    hist_info(n_cl, pred, gt)                 -> (hist[n_cl, n_cl], labeled, correct)       metric.py:10-18
    compute_metric(results, per_class=False)  -> (mean_IU, mean_pixel_acc[, iu, class_acc]) metric.py:21-39
    compute_score(hist, correct, labeled)                                                   metric.py:42-49
    compute_score_per_class(hist, correct, labeled)                                         metric.py:51-64

The per-pixel work -- the confusion histogram over the whole dataset -- runs on the GPU (integer,
bit-exact; ``mss_confusion_hist`` / ``mss_confusion_from_logits``).  What is left for the host is the
closed-form arithmetic on the 19 x 19 matrix; it sits behind the C ABI as well (``mss_confusion_scores``, host-only,
IEEE float64 operation by operation in numpy's order -- pinned to reference-generated fixtures by the CPU test-suite),
so C callers get the same numbers.  ``ConfusionAccumulator`` is the streaming form: it keeps
``hist/labeled/correct`` on the device across batches (``compute_metric``'s accumulation loop) and can take
the NCHW logits directly (argmax fused, the int64 prediction map is never written).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L

__all__ = ["hist_info", "compute_metric", "compute_score", "compute_score_per_class", "ConfusionAccumulator"]


def _as_index_map(x, device) -> torch.Tensor:
    t = torch.as_tensor(x)
    if t.dtype == torch.bool:
        t = t.to(torch.uint8)
    if t.dtype.is_floating_point:
        raise TypeError("class-index maps must be integer tensors")
    if t.dtype not in (torch.uint8, torch.int32, torch.int64):
        t = t.to(torch.int64)
    return t.to(device=device, non_blocking=True).reshape(-1).contiguous()


def _device(*xs) -> torch.device:
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    if not torch.cuda.is_available():
        raise L.MssError("no CUDA device: multishiftseg_b200 computes this path on the GPU only")
    return torch.device("cuda", torch.cuda.current_device())


class ConfusionAccumulator:
    """Device-resident ``hist`` / ``labeled`` / ``correct`` of metric.py:21-33, updated per batch."""

    def __init__(self, n_cl: int = 19, device=None):
        if not 1 <= int(n_cl) <= 32:
            raise ValueError("n_cl must be in 1..32")
        self.n_cl = int(n_cl)
        self.device = _device() if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise L.MssError("ConfusionAccumulator needs a CUDA device (no CPU fallback)")
        L.load()
        self.hist = torch.zeros(self.n_cl * self.n_cl, dtype=torch.int64, device=self.device)
        self.lc = torch.zeros(3, dtype=torch.int64, device=self.device)

    def reset(self):
        self.hist.zero_()
        self.lc.zero_()

    def update(self, pred, gt):
        """pred, gt: class-index maps of equal shape (metric.py:11 asserts the same)."""
        if tuple(np.shape(pred)) != tuple(np.shape(gt)):
            raise AssertionError("pred.shape != gt.shape")               # metric.py:11
        p = _as_index_map(pred, self.device)
        g = _as_index_map(gt, self.device)
        with torch.cuda.device(self.device):
            L.check(L.load().mss_confusion_hist(p.data_ptr(), L.label_code(p), g.data_ptr(), L.label_code(g),
                                                p.numel(), self.n_cl, self.hist.data_ptr(), self.lc.data_ptr(),
                                                L.stream_ptr(self.device)), "mss_confusion_hist")

    def update_from_logits(self, logits: torch.Tensor, gt):
        """logits [B, C, H, W] fp32 on the GPU; pred = logits.argmax(1) is fused into the histogram pass."""
        L.require_cuda(logits, "logits")
        if logits.dim() != 4 or logits.dtype != torch.float32:
            raise TypeError("logits must be a [B, C, H, W] float32 tensor")
        B, Cc, H, W = logits.shape
        if tuple(np.shape(gt)) != (B, H, W):
            raise AssertionError("gt must have shape [B, H, W]")
        x = logits.contiguous()
        g = _as_index_map(gt, self.device)
        with torch.cuda.device(self.device):
            L.check(L.load().mss_confusion_from_logits(x.data_ptr(), B, Cc, H * W, g.data_ptr(), L.label_code(g),
                                                       self.n_cl, self.hist.data_ptr(), self.lc.data_ptr(),
                                                       L.stream_ptr(self.device)), "mss_confusion_from_logits")

    def result(self):
        """-> (hist [n_cl, n_cl] int64 ndarray, labeled, correct) -- the one synchronisation of a streaming evaluation.
        Raises ``ValueError`` if any update saw a labeled pixel whose ``n_cl * gt + pred`` is out of range (numpy's
        bincount / reshape raise in the reference, metric.py:15-17)."""
        import ctypes as C
        h = np.empty(self.n_cl * self.n_cl, dtype=np.int64)
        lc = (C.c_int64 * 2)()
        with torch.cuda.device(self.device):
            L.check(L.load().mss_confusion_result(self.hist.data_ptr(), self.lc.data_ptr(), self.n_cl, h.ctypes.data, lc,
                                                  L.stream_ptr(self.device)), "mss_confusion_result")
        return h.reshape(self.n_cl, self.n_cl), np.int64(lc[0]), np.int64(lc[1])

    def compute(self, per_class: bool = False):
        h, labeled, correct = self.result()
        return compute_metric([{"hist": h, "labeled": labeled, "correct": correct}], per_class=per_class)


def hist_info(n_cl, pred, gt):
    """metric.py:10-18.  Returns ``(hist, labeled, correct)`` with ``hist`` an ``[n_cl, n_cl]`` int64 array."""
    acc = ConfusionAccumulator(n_cl, _device(pred, gt))
    acc.update(pred, gt)
    return acc.result()


def _scores(hist, correct, labeled, per_class: bool):
    """metric.py:42-64 through the C ABI (``mss_confusion_scores``; host-only, no device work)."""
    import ctypes as C
    import warnings
    h = np.ascontiguousarray(hist, dtype=np.float64)
    n_cl = h.shape[0]
    if h.shape != (n_cl, n_cl):
        raise ValueError("hist must be a square [n_cl, n_cl] matrix")
    iu = np.empty(n_cl, dtype=np.float64)
    acc = np.empty(n_cl, dtype=np.float64)
    out = (C.c_double * 3)()
    L.check(L.load().mss_confusion_scores(h.ctypes.data, n_cl, float(correct), float(labeled), 1 if per_class else 0,
                                          iu.ctypes.data, acc.ctypes.data, out), "mss_confusion_scores")
    if np.isnan(iu).any() or float(labeled) == 0.0:          # the reference's numpy expressions warn here (0 / 0)
        warnings.warn("invalid value encountered in divide", RuntimeWarning, stacklevel=3)
    return iu, acc, np.float64(out[0]), np.float64(out[1]), np.float64(out[2])


def compute_score(hist, correct, labeled):
    """metric.py:42-49 -> (iu, mean_IU, mean_IU_no_back, mean_pixel_acc); classes absent from both maps give nan."""
    iu, _, mean_iu, mean_iu_nb, pix = _scores(hist, correct, labeled, False)
    return iu, mean_iu, mean_iu_nb, pix


def compute_score_per_class(hist, correct, labeled):
    """metric.py:51-64 -> (iu, mean_IU, class_acc, mean_pixel_acc)."""
    iu, acc, mean_iu, _, pix = _scores(hist, correct, labeled, True)
    return iu, mean_iu, acc, pix


def compute_metric(results, per_class=False):
    """metric.py:21-39: sum the per-batch dicts (the reference hard-codes a 19 x 19 float64 accumulator)."""
    hist = np.zeros((19, 19))
    correct = 0
    labeled = 0
    for d in results:
        hist += d["hist"]
        correct += d["correct"]
        labeled += d["labeled"]
    if per_class:
        iu, mean_IU, class_acc, mean_pixel_acc = compute_score_per_class(hist, correct, labeled)
        return mean_IU, mean_pixel_acc, iu, class_acc
    iu, mean_IU, _, mean_pixel_acc = compute_score(hist, correct, labeled)
    return mean_IU, mean_pixel_acc
