"""CPU oracle (numpy) for the mIoU half of the reference's ``lib/utils/metric.py``.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Parity status: PINNED -- ``tests/golden/make_golden_segmetric.py`` executes the reference's own
``hist_info`` / ``compute_metric`` (the module loads by file path) and ``tests/golden/segmetric_golden.json``
holds the outputs (integers verbatim, float64 as hex).  Every function cites the reference lines it follows
(relative to /root/reference).
"""
from __future__ import annotations

import numpy as np


def hist_info(n_cl: int, pred: np.ndarray, gt: np.ndarray):
    """lib/utils/metric.py:10-18 -- (hist [n_cl, n_cl], labeled, correct).

    Restated with a plain counting loop over the labeled pixels' flat index ``n_cl * gt + pred`` (what
    ``np.bincount(..., minlength=n_cl**2)`` counts), not by calling the same numpy one-liner."""
    pred = np.asarray(pred).reshape(-1).astype(np.int64)
    gt = np.asarray(gt).reshape(-1).astype(np.int64)
    assert pred.shape == gt.shape                                       # :11
    keep = (gt >= 0) & (gt < n_cl)                                      # :12
    labeled = int(keep.sum())                                           # :13
    correct = int((pred[keep] == gt[keep]).sum())                       # :14
    flat = n_cl * gt[keep] + pred[keep]                                 # :16
    if flat.size and (flat.min() < 0 or flat.max() >= n_cl * n_cl):
        raise ValueError("n_cl*gt + pred outside [0, n_cl^2): numpy's bincount / reshape raises")
    hist = np.zeros(n_cl * n_cl, dtype=np.int64)
    vals, counts = np.unique(flat, return_counts=True)
    hist[vals] = counts
    return hist.reshape(n_cl, n_cl), labeled, correct


def argmax_first(logits: np.ndarray) -> np.ndarray:
    """``logit.argmax(1)`` with torch's tie rule (index of the first maximal value) -- what a caller of
    hist_info passes for a ``[B, C, H, W]`` logit map."""
    return np.argmax(logits, axis=1)                                    # numpy: first occurrence too


def compute_score(hist, correct, labeled):
    """lib/utils/metric.py:42-49."""
    hist = np.asarray(hist, dtype=np.float64)
    diag = np.diag(hist)
    with np.errstate(divide="ignore", invalid="ignore"):
        iu = diag / (hist.sum(1) + hist.sum(0) - diag)                  # :43
        mean_IU = np.nanmean(iu)                                        # :44
        mean_IU_no_back = np.nanmean(iu[1:])                            # :45
        mean_pixel_acc = correct / labeled                              # :48
    return iu, mean_IU, mean_IU_no_back, mean_pixel_acc


def compute_score_per_class(hist, correct, labeled):
    """lib/utils/metric.py:51-64."""
    hist = np.asarray(hist, dtype=np.float64)
    inter = np.diag(hist)
    union = hist.sum(axis=1) + hist.sum(axis=0) - inter
    iu = inter / np.maximum(union, 1)                                   # :56
    class_acc = inter / np.maximum(hist.sum(axis=1), 1)                 # :59
    return iu, np.nanmean(iu), class_acc, correct / labeled             # :62-64


def compute_metric(results, per_class=False):
    """lib/utils/metric.py:21-39 (19 x 19 float64 accumulator, as hard-coded there)."""
    hist = np.zeros((19, 19))
    correct = labeled = 0
    for d in results:
        hist += d["hist"]
        correct += d["correct"]
        labeled += d["labeled"]
    if per_class:
        iu, mean_IU, class_acc, acc = compute_score_per_class(hist, correct, labeled)
        return mean_IU, acc, iu, class_acc
    _, mean_IU, _, acc = compute_score(hist, correct, labeled)
    return mean_IU, acc
