"""CPU stand-in for evaluator.CudaBackend, built on the oracle (tests only): lets the gloo world_size-2
tests exercise the host / collective logic of StreamingEvaluator without a GPU."""
import numpy as np
import torch

from oracle import c_oracle, metrics_oracle as mo


class _Buf:
    def __init__(self, capacity):
        self.capacity = capacity
        self.reset()

    def reset(self):
        self.keys = np.zeros(0, np.uint32)
        self.labs = np.zeros(0, np.uint8)
        self.nan = self.inf = 0


class NumpyBackend:
    def new_buffer(self, capacity):
        return _Buf(capacity)

    def append(self, buf, scores, labels, id_in, id_out):
        s = np.asarray(scores, dtype=np.float32).ravel()
        l = np.asarray(labels).ravel()
        v = (l == id_in) | (l == id_out)
        buf.nan |= int(np.isnan(s[v]).any())
        buf.inf |= int(np.isinf(s[v]).any())
        buf.keys = np.concatenate([buf.keys, mo.float_key_desc(s[v])])
        buf.labs = np.concatenate([buf.labs, (l[v] == id_out).astype(np.uint8)])
        assert buf.keys.size <= buf.capacity

    def state(self, buf):
        return buf.keys.size, int(buf.labs.sum()), buf.nan, buf.inf

    def pairs(self, buf, m):
        return torch.from_numpy(buf.keys[:m].view(np.int32).copy()), torch.from_numpy(buf.labs[:m].copy())

    @staticmethod
    def _u32(t):
        return t.numpy().view(np.uint32)

    def histogram(self, keys, m, bits, every=1):
        k = self._u32(keys)[:m]
        if every > 1:                                   # systematic sample of 4-key groups, like the CUDA kernel
            k = k[: m - m % 4].reshape(-1, 4)[::every].reshape(-1)
        return torch.from_numpy(np.bincount(k >> np.uint32(32 - bits), minlength=1 << bits).astype(np.int64))

    def partition(self, keys, labs, m, splitters, parts):
        k = self._u32(keys)[:m]
        dest = np.searchsorted(np.asarray(splitters, dtype=np.uint64), k.astype(np.uint64), side="right")
        order = np.argsort(dest, kind="stable")
        counts = np.bincount(dest, minlength=parts).tolist()
        return (torch.from_numpy(k[order].view(np.int32).copy()), torch.from_numpy(labs.numpy()[:m][order].copy()),
                [int(c) for c in counts])

    def sort(self, keys, labs, m):
        k = self._u32(keys)[:m]
        order = np.argsort(k, kind="stable")
        keys[:m] = torch.from_numpy(k[order].view(np.int32).copy())
        labs[:m] = torch.from_numpy(labs.numpy()[:m][order].copy())

    def counts(self, keys, labs, m, pos_before, idx_before):
        k = self._u32(keys)[:m]
        y = labs.numpy()[:m].astype(np.int64)
        ends = np.r_[np.nonzero(k[1:] != k[:-1])[0], m - 1]
        tps = np.cumsum(y)[ends] + pos_before
        fps = ends + 1 + idx_before - tps
        return torch.from_numpy(tps.astype(np.int64)), torch.from_numpy(fps.astype(np.int64))

    def counts_local(self, keys, labs, m):
        tps, fps = self.counts(keys, labs, m, 0, 0)
        return tps, fps, int(labs.numpy()[:m].astype(np.int64).sum())

    def tail(self, tps, fps, recall_level=0.95):
        return tuple(np.float64(v) for v in c_oracle.metrics_from_counts(tps.numpy(), fps.numpy(), recall_level))

    def tensor(self, data, dtype):
        return torch.tensor(data, dtype=dtype)

    def empty(self, n, dtype):
        return torch.empty(n, dtype=dtype)
