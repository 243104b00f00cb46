"""CPU oracle for the OOD metric stage (AUROC / AP / FPR@95TPR).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Parity status: PINNED.  ``tests/golden/make_golden.py`` executes the reference's
own ``lib/utils/metric.py`` (loaded by file path from /root/reference) on seeded
inputs and commits the outputs as hex float64 under ``tests/golden/``;
``tests/test_oracle_metrics.py`` checks every function here against those
vectors and against the SURVEY.md section 8(c) known-answer tests K1-K13.

Two independent restatements live here:

* ``eval_ood_measure`` / ``get_measures`` / ``fpr_and_fdr_at_recall`` --
  the reference's control flow (``lib/utils/metric.py:69-180``) with the same
  third-party call sites: ``sklearn.metrics.roc_auc_score`` (metric.py:142) and
  ``sklearn.metrics.average_precision_score`` (metric.py:146).  scikit-learn is
  the dependency the arithmetic lives in; the reference pins scikit-learn==1.4.0
  (environment.yml:332), this image ships 1.9.0 and parity is defined against
  the sklearn in the image (same one the GPU box has).

* ``ood_counts`` + ``metrics_from_counts`` -- the integer-count specification
  (SURVEY.md section 8(c) "Exact metric specification"), which is what the CUDA
  path implements: per distinct float32 threshold the int64 pair (tps, fps),
  then a float64 tail that replays numpy's pairwise summation tree
  (``pairwise_sum``).  No sklearn involved; ``==`` to the first restatement.
"""
from __future__ import annotations

import numpy as np

PW_BLOCKSIZE = 128  # numpy/_core/src/umath/loops_utils.h.src


# ----------------------------------------------------------------------------
# (1) reference-flow restatement (metric.py:69-180), sklearn at the call sites
# ----------------------------------------------------------------------------
def stable_cumsum(arr, rtol=1e-05, atol=1e-08):
    """metric.py:69-85 -- float64 cumsum whose last element is checked against sum."""
    out = np.cumsum(arr, dtype=np.float64)
    expected = np.sum(arr, dtype=np.float64)
    if not np.allclose(out[-1], expected, rtol=rtol, atol=atol):
        raise RuntimeError("cumsum was found to be unstable: "
                           "its last element does not correspond to sum")
    return out


def fpr_and_fdr_at_recall(y_true, y_score, recall_level=0.95, pos_label=None):
    """metric.py:87-127 -- FPR at the threshold whose recall is closest to 0.95."""
    classes = np.unique(y_true)
    if pos_label is None and not any(
            np.array_equal(classes, c) for c in ([0, 1], [-1, 1], [0], [-1], [1])):
        raise ValueError("Data is not binary and pos_label is not specified")
    elif pos_label is None:
        pos_label = 1.0
    y_true = (y_true == pos_label)

    order = np.argsort(y_score, kind="mergesort")[::-1]       # metric.py:103
    y_score = y_score[order]
    y_true = y_true[order]

    distinct = np.where(np.diff(y_score))[0]                  # metric.py:110
    threshold_idxs = np.r_[distinct, y_true.size - 1]

    tps = stable_cumsum(y_true)[threshold_idxs]               # metric.py:114
    fps = 1 + threshold_idxs - tps

    recall = tps / tps[-1]
    last_ind = tps.searchsorted(tps[-1])                      # metric.py:121
    sl = slice(last_ind, None, -1)
    recall, fps = np.r_[recall[sl], 1], np.r_[fps[sl], 0]

    cutoff = np.argmin(np.abs(recall - recall_level))         # metric.py:125
    return fps[cutoff] / (np.sum(np.logical_not(y_true)))


def get_measures(_pos, _neg, recall_level=0.95):
    """metric.py:130-153."""
    import sklearn.metrics as sk
    pos = np.array(_pos[:]).reshape((-1, 1))
    neg = np.array(_neg[:]).reshape((-1, 1))
    examples = np.squeeze(np.vstack((pos, neg)))
    labels = np.zeros(len(examples), dtype=np.int32)
    labels[:len(pos)] += 1
    auroc = sk.roc_auc_score(labels, examples)                # metric.py:142
    aupr = sk.average_precision_score(labels, examples)       # metric.py:146
    fpr = fpr_and_fdr_at_recall(labels, examples, recall_level)
    return auroc, aupr, fpr


def eval_ood_measure(conf, seg_label, train_id_in=0, train_id_out=1):
    """metric.py:170-180 (get_and_print_results' np.mean of 1-element lists is a no-op)."""
    in_scores = conf[seg_label == train_id_in]
    out_scores = conf[seg_label == train_id_out]
    if (len(out_scores) != 0) and (len(in_scores) != 0):
        auroc, aupr, fpr = get_measures(out_scores, in_scores)
        return np.mean([auroc]), np.mean([aupr]), np.mean([fpr])
    return None


# ----------------------------------------------------------------------------
# (2) integer-count specification + float64 tail (what the CUDA path implements)
# ----------------------------------------------------------------------------
def float_key_desc(scores: np.ndarray) -> np.ndarray:
    """uint32 key whose ASCENDING order is the DESCENDING order of the float32
    score; -0.0 and +0.0 share one key (np.diff(y_score) == 0 for them,
    metric.py:110 / sklearn _ranking.py:916)."""
    s = np.ascontiguousarray(scores, dtype=np.float32).ravel()
    u = s.view(np.uint32).copy()
    u[u == np.uint32(0x80000000)] = 0                          # -0.0 -> +0.0
    neg = (u >> np.uint32(31)).astype(bool)
    asc = np.where(neg, ~u, u | np.uint32(0x80000000))         # ascending-order key
    return ~asc


def key_to_float(key: np.ndarray) -> np.ndarray:
    asc = ~np.asarray(key, dtype=np.uint32)
    neg = (asc >> np.uint32(31)) == 0
    u = np.where(neg, ~asc, asc & np.uint32(0x7FFFFFFF))
    return u.view(np.float32)


def ood_counts(conf, seg_label, train_id_in=0, train_id_out=1):
    """Per distinct float32 threshold (descending score): cumulative
    (tps, fps) as int64.  Returns (tps, fps) or None when a class is empty
    (metric.py:176-180).  Raises ValueError on non-finite valid scores like
    sklearn's assert_all_finite (_ranking.py:896-897)."""
    conf = np.asarray(conf).ravel()
    lab = np.asarray(seg_label).ravel()
    valid = (lab == train_id_in) | (lab == train_id_out)
    s = conf[valid].astype(np.float32, copy=False)
    y = (lab[valid] == train_id_out)
    if y.sum() == 0 or (~y).sum() == 0:
        return None
    if np.isnan(s).any():
        raise ValueError("Input contains NaN.")
    if np.isinf(s).any():
        raise ValueError("Input contains infinity or a value too large for dtype('float32').")
    key = float_key_desc(s)
    order = np.argsort(key, kind="stable")
    key = key[order]
    y = y[order]
    ends = np.r_[np.nonzero(key[1:] != key[:-1])[0], key.size - 1]
    tps = np.cumsum(y, dtype=np.int64)[ends]
    fps = ends.astype(np.int64) + 1 - tps
    return tps, fps


def pairwise_sum(a: np.ndarray) -> float:
    """numpy's float64 pairwise summation (numpy/_core/src/umath/loops_utils.h.src,
    ``@TYPE@_pairwise_sum``): n<8 sequential; n<=128 eight interleaved
    accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the n%8
    tail sequentially; else split at n/2 rounded down to a multiple of 8."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.size
    if n < 8:
        res = np.float64(-0.0)          # numpy starts from -0.0 to preserve -0
        for i in range(n):
            res = res + a[i]
        return float(res)
    if n <= PW_BLOCKSIZE:
        r = a[:8].copy()
        i = 8
        while i < n - (n % 8):
            r += a[i:i + 8]             # eight independent accumulators
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + a[i]
            i += 1
        return float(res)
    n2 = n // 2
    n2 -= n2 % 8
    return float(np.float64(pairwise_sum(a[:n2])) + np.float64(pairwise_sum(a[n2:])))


def pairwise_leaves(n: int):
    """Leaf (start, length) list, left to right, of numpy's pairwise tree over n terms."""
    out = []
    stack = [(0, n)]
    while stack:
        s, m = stack.pop()
        if m <= PW_BLOCKSIZE:
            out.append((s, m))
        else:
            n2 = m // 2
            n2 -= n2 % 8
            stack.append((s + n2, m - n2))
            stack.append((s, n2))
    return out


def metrics_from_counts(tps: np.ndarray, fps: np.ndarray, recall_level=0.95):
    """float64 tail over int64 (tps, fps); follows sklearn _ranking.py:1331-1378
    (roc_curve, drop_intermediate=True), :53-116 (auc) + scipy trapezoid
    (_quadrature.py:153-156), :1160-1208 (precision_recall_curve), :243-260 (AP)
    and metric.py:116-127 (FPR@95).  Sums use ``pairwise_sum`` above, not np.sum."""
    tps = np.asarray(tps, dtype=np.int64)
    fps = np.asarray(fps, dtype=np.int64)
    T = tps.size
    P = np.float64(tps[-1])
    N = np.float64(fps[-1])
    tf = tps.astype(np.float64)
    ff = fps.astype(np.float64)

    # ---- AUROC
    if T > 2:
        d2f = fps[2:] - 2 * fps[1:-1] + fps[:-2]
        d2t = tps[2:] - 2 * tps[1:-1] + tps[:-2]
        keep = np.r_[True, (d2f != 0) | (d2t != 0), True]
    else:
        keep = np.ones(T, dtype=bool)
    fpr = np.r_[0.0, ff[keep]] / N
    tpr = np.r_[0.0, tf[keep]] / P
    terms = (fpr[1:] - fpr[:-1]) * (tpr[1:] + tpr[:-1]) / 2.0
    auroc = pairwise_sum(terms)

    # ---- AP
    prec = tf / (tf + ff)
    rec = tf / P
    prec_r = np.r_[prec[::-1], 1.0]
    rec_r = np.r_[rec[::-1], 0.0]
    ap_terms = (rec_r[1:] - rec_r[:-1]) * prec_r[:-1]
    ap = max(0.0, -pairwise_sum(ap_terms))

    # ---- FPR@95
    last = int(np.searchsorted(tps, tps[-1]))
    d = np.abs(rec[: last + 1] - recall_level)
    m = d.min()
    k = int(np.nonzero(d == m)[0][-1])          # ties -> largest original index
    fpr95 = float(ff[k] / N)
    return float(auroc), float(ap), fpr95


def eval_ood_measure_counts(conf, seg_label, train_id_in=0, train_id_out=1):
    c = ood_counts(conf, seg_label, train_id_in, train_id_out)
    if c is None:
        return None
    return metrics_from_counts(*c)
