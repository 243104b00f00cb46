"""CPU stand-in for evaluator.CudaBackend INCLUDING peer-mapped buffers (tests only): POSIX shared memory plays the
peer-mapped receive buffers, a file lock the system-scope atomicAdd, memmove the copy engines.  With the CUDA stream /
event calls of the evaluator stubbed out (`install_cuda_stubs`), the gloo world_size 2 / 3 tests run the host logic of the
STREAMED exchange -- staging, copy plans, double-buffered outboxes, reservations, flush order, reset / reuse -- without
a GPU.  Nothing here is imported by the product."""
import ctypes as C
import fcntl
import os
from multiprocessing import shared_memory

import numpy as np
import torch

from multishiftseg_b200 import _lib as L
from numpy_backend import NumpyBackend, _t, _u32  # noqa: F401
from oracle import metrics_oracle as mo

STATE = L.EVAL_STATE_BYTES          # {n_neg u64, n_pos u64, nan u32, inf u32, overflow u64, pad}


def _u64_at(addr):
    return C.c_uint64.from_address(addr)


def _keys_at(addr, n):
    return np.ctypeslib.as_array((C.c_uint32 * n).from_address(addr))


class PeerBuf:
    """Two-ended evaluator buffer over caller-provided memory (keys tensor + 64-byte state tensor)."""

    def __init__(self, keys: torch.Tensor, state: torch.Tensor, capacity: int):
        self.keys, self.state, self.capacity = keys, state, int(capacity)

    def _s(self):
        return self.state.numpy()

    def reset(self):
        self.state.zero_()

    def counts(self):
        s = self._s()
        n_neg, n_pos = (int(v) for v in s[:16].view(np.uint64))
        over = int(s[24:32].view(np.uint64)[0])
        nan, inf = (int(v) for v in s[16:24].view(np.uint32))
        return n_neg, n_pos, nan, inf, over

    def read_state(self):
        n_neg, n_pos, nan, inf, over = self.counts()
        if over or n_neg + n_pos > self.capacity:
            raise L.MssError(f"evaluator capacity {self.capacity} exceeded")
        return n_neg + n_pos, n_pos, nan, inf

    def streams(self, m, n_pos):
        return self.keys[: m - n_pos], self.keys[self.capacity - n_pos: self.capacity]


class FakePeerBackend(NumpyBackend):
    def __init__(self, lock_path: str):
        self.device = torch.device("cpu")
        self.lock_path = lock_path
        self._keep, self._own = [], []
        self.copies = 0

    # -- buffers: real two-ended memory with a state, like the product's PairBuffer
    def new_buffer(self, capacity):
        cap = max(int(capacity), 1)
        b = PeerBuf(torch.zeros(cap, dtype=torch.int32), torch.zeros(STATE, dtype=torch.uint8), cap)
        return b

    def append(self, buf, scores, labels, id_in, id_out):
        s = np.asarray(scores, dtype=np.float32).ravel()
        l = np.asarray(labels).ravel()
        neg, pos = mo.float_key_desc(s[l == id_in]), mo.float_key_desc(s[l == id_out])
        v = (l == id_in) | (l == id_out)
        st = buf.state.numpy()
        cnt = st[:16].view(np.uint64)
        flags = st[16:24].view(np.uint32)
        k = buf.keys.numpy().view(np.uint32)
        a, b = int(cnt[0]), int(cnt[1])
        assert a + neg.size + b + pos.size <= buf.capacity
        k[a: a + neg.size] = neg
        k[buf.capacity - b - pos.size: buf.capacity - b] = pos[::-1]                # positives grow downwards
        cnt[0], cnt[1] = a + neg.size, b + pos.size
        flags[0] |= int(np.isnan(s[v]).any())
        flags[1] |= int(np.isinf(s[v]).any())

    def state(self, buf):
        return buf.read_state()

    def streams(self, buf, m, n_pos):
        return buf.streams(m, n_pos)

    # -- "peer-mapped" memory
    def peer_alloc(self, capacity: int):
        kb = (int(capacity) * 4 + 255) // 256 * 256
        shm = shared_memory.SharedMemory(create=True, size=kb + STATE)
        self._keep.append(shm)
        self._own.append(shm)
        return shm

    def peer_map(self, raw, capacity: int, group):
        import torch.distributed as dist
        cap = int(capacity)
        kb = (cap * 4 + 255) // 256 * 256
        names = [None] * dist.get_world_size(group)
        dist.all_gather_object(names, raw.name, group=group)
        segs = [raw if n == raw.name else shared_memory.SharedMemory(name=n) for n in names]
        self._keep.extend(segs)
        ptrs = [C.addressof(C.c_char.from_buffer(s.buf)) for s in segs]
        mine = torch.frombuffer(raw.buf, dtype=torch.uint8)
        keys, state = mine[: cap * 4].view(torch.int32), mine[kb: kb + STATE]
        return {"cap": cap, "raw": raw, "keys": keys, "buf": PeerBuf(keys, state, cap), "key_ptrs": ptrs,
                "state_ptrs": [p + kb for p in ptrs]}

    def _atomic_add(self, addr: int, v: int) -> int:
        with open(self.lock_path, "a") as f:
            fcntl.flock(f, fcntl.LOCK_EX)
            cell = _u64_at(addr)
            old = int(cell.value)
            cell.value = old + int(v)
            fcntl.flock(f, fcntl.LOCK_UN)
        return old

    # -- the staged exchange (mss_eval_exchange_stage + mss_memcpy_async in the product)
    def exchange_stage(self, buf, spl_dev, parts, out_key_ptrs, out_state_ptrs, out_capacity, recv_state_ptrs, recv_capacity,
                       accum_state, plan):
        n_neg, n_pos, nan, inf, over = buf.counts()
        k = buf.keys.numpy().view(np.uint32)
        spl = spl_dev.numpy().view(np.uint32)[: parts - 1].astype(np.uint64)
        for strm, arr in ((0, k[:n_neg]), (1, k[buf.capacity - n_pos: buf.capacity])):
            dest = np.searchsorted(spl, arr.astype(np.uint64), side="right")
            for d in range(parts):
                run = arr[dest == d]
                cnt, off = int(run.size), 0
                box = _keys_at(out_key_ptrs[d], out_capacity)
                if strm == 0:
                    box[:cnt] = run
                else:
                    box[out_capacity - cnt: out_capacity] = run
                if cnt:
                    base = self._atomic_add(recv_state_ptrs[d] + 8 * strm, cnt)
                    if base + cnt > recv_capacity:
                        self._atomic_add(recv_state_ptrs[d] + 24, cnt)
                        cnt = 0
                    else:
                        off = recv_capacity - base - cnt if strm else base
                plan[4 * d + 2 * strm] = cnt
                plan[4 * d + 2 * strm + 1] = off
        a = accum_state.numpy()
        a[:16].view(np.uint64)[:] += np.array([n_neg, n_pos], dtype=np.uint64)
        a[16:24].view(np.uint32)[:] |= np.array([nan, inf], dtype=np.uint32)
        a[24:32].view(np.uint64)[0] += np.uint64(over)
        buf.reset()

    def memcpy_async(self, dst_ptr, src_ptr, nbytes, stream):
        C.memmove(int(dst_ptr), int(src_ptr), int(nbytes))
        self.copies += 1

    def close(self):
        """Call after a barrier: nobody uses the segments any more.  (Tensors may still view the buffers: only unlink.)"""
        for s in self._own:
            try:
                s.unlink()
            except Exception:
                pass


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_event(self, ev):
        pass

    def wait_stream(self, s):
        pass

    def synchronize(self):
        pass


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass


def install_cuda_stubs():
    """The evaluator's stream / event calls become no-ops (everything the fake backend does is synchronous)."""
    cur = _Stream()
    torch.cuda.current_stream = lambda device=None: cur
    torch.cuda.Stream = _Stream
    torch.cuda.Event = _Event
    torch.cuda.synchronize = lambda device=None: None
    torch.Tensor.pin_memory = lambda self, *a, **k: self
