"""CPU stand-in for evaluator.CudaBackend, built on the oracle (tests only): lets the gloo world_size-2
tests exercise the host / collective logic of StreamingEvaluator without a GPU.  Same two-stream layout as the
product (in-distribution keys and OOD keys kept apart, no label bytes)."""
import numpy as np
import torch

from oracle import c_oracle, metrics_oracle as mo


class _Buf:
    def __init__(self, capacity):
        self.capacity = capacity
        self.reset()

    def reset(self):
        self.neg = np.zeros(0, np.uint32)
        self.pos = np.zeros(0, np.uint32)
        self.nan = self.inf = 0


def _t(k):
    return torch.from_numpy(np.ascontiguousarray(k).view(np.int32).copy())


def _u32(t, n):
    return t.numpy().view(np.uint32)[:n]


def merged_counts(neg, pos, pos_before=0, neg_before=0):
    """sorted uint32 streams -> (tps, fps) per distinct key of the union (the integer spec of SURVEY 8c)."""
    keys = np.unique(np.concatenate([neg, pos]))
    tps = np.searchsorted(pos, keys, side="right").astype(np.int64) + pos_before
    fps = np.searchsorted(neg, keys, side="right").astype(np.int64) + neg_before
    return tps, fps


class NumpyBackend:
    def new_buffer(self, capacity):
        return _Buf(capacity)

    def append(self, buf, scores, labels, id_in, id_out):
        s = np.asarray(scores, dtype=np.float32).ravel()
        l = np.asarray(labels).ravel()
        v = (l == id_in) | (l == id_out)
        buf.nan |= int(np.isnan(s[v]).any())
        buf.inf |= int(np.isinf(s[v]).any())
        buf.neg = np.concatenate([buf.neg, mo.float_key_desc(s[l == id_in])])
        buf.pos = np.concatenate([buf.pos, mo.float_key_desc(s[l == id_out])])
        assert buf.neg.size + buf.pos.size <= buf.capacity

    def state(self, buf):
        return buf.neg.size + buf.pos.size, buf.pos.size, buf.nan, buf.inf

    def streams(self, buf, m, n_pos):
        return _t(buf.neg), _t(buf.pos)

    def finish_local(self, buf, recall_level):
        m, n_pos, nan, inf = self.state(buf)
        if n_pos == 0 or n_pos == m:
            return None
        if nan:
            raise ValueError("Input contains NaN.")
        if inf:
            raise ValueError("Input contains infinity or a value too large for dtype('float32').")
        tps, fps = merged_counts(np.sort(buf.neg), np.sort(buf.pos))
        return self.tail(torch.from_numpy(tps), torch.from_numpy(fps), recall_level)

    def histogram(self, keys, m, bits, every=1):
        k = _u32(keys, m)
        if every > 1:                                   # systematic sample of 4-key groups, like the CUDA kernel
            k = k[: m - m % 4].reshape(-1, 4)[::every].reshape(-1)
        return torch.from_numpy(np.bincount(k >> np.uint32(32 - bits), minlength=1 << bits).astype(np.int64))

    def histogram_refine(self, keys, m, prefix16, every=1):
        k = _u32(keys, m)
        if every > 1:
            k = k[: m - m % 4].reshape(-1, 4)[::every].reshape(-1)
        k = k[(k >> np.uint32(16)) == np.uint32(prefix16)]
        return torch.from_numpy(np.bincount(k & np.uint32(0xFFFF), minlength=1 << 16).astype(np.int64))

    def partition(self, keys, m, splitters, parts):
        k = _u32(keys, m)
        dest = np.searchsorted(np.asarray(splitters, dtype=np.uint64), k.astype(np.uint64), side="right")
        order = np.argsort(dest, kind="stable")
        counts = np.bincount(dest, minlength=parts).tolist()
        return _t(k[order]), [int(c) for c in counts]

    def sort2(self, neg, n_neg, pos, n_pos):
        for t, n in ((neg, n_neg), (pos, n_pos)):
            if n:
                t[:n] = _t(np.sort(_u32(t, n)))

    def counts(self, neg, n_neg, pos, n_pos, pos_before, neg_before):
        tps, fps = merged_counts(_u32(neg, n_neg), _u32(pos, n_pos), pos_before, neg_before)
        return torch.from_numpy(tps), torch.from_numpy(fps)

    def tail(self, tps, fps, recall_level=0.95):
        return tuple(np.float64(v) for v in c_oracle.metrics_from_counts(tps.numpy(), fps.numpy(), recall_level))

    def tensor(self, data, dtype):
        return torch.tensor(data, dtype=dtype)

    def empty(self, n, dtype):
        return torch.empty(n, dtype=dtype)
