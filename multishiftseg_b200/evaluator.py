"""Streaming, multi-GPU evaluator: the on-device replacement of the reference's tester loop
(test_deeplab.py:84-117, test_m2f.py:125-158, and ``valid_batch`` in both trainers), i.e.

    anomaly_scores.append(score.cpu().numpy()); ood_gts.append(target.cpu().numpy())   # per batch
    eval_ood_measure(np.concatenate(anomaly_scores), np.concatenate(ood_gts))           # at the end

``update`` appends the valid pixels of a batch to a device buffer of order-preserving keys (no D2H);
``compute`` returns exactly what ``eval_ood_measure`` returns for the concatenated dataset.

Multi-GPU (one process per GPU, images sharded over ranks, ``torch.distributed``): the only exchange
step of the whole path.  Integer-exact, so the N-GPU result is bit-identical to the 1-GPU result:

  1. all_reduce of (count, positives, nan, inf)                          -> None / ValueError decisions
  2. all_reduce of a 2^16-bin histogram of the top key bits (+ a second level inside heavy bins)
                                                                         -> identical splitters everywhere
  3. exchange by key range: rank r owns key range r, a distinct score never straddles two ranks.  The keys travel
     as two key-only streams (in-distribution / OOD; the label is the stream) into the owner's peer-mapped buffer:
     "stream" (per batch, behind the scoring kernel: local partition + copy engines), "stream_sm", "p2p",
     "p2p_counted", or "nccl" (partition + all_to_all) -- see StreamingEvaluator.__init__ and DESIGN.md section 5
  4. local radix sort of both streams, merge-path counts with the global (negatives, positives) prefixes of the
     lower ranks, known from the all-gathered stream sizes
  5. all_gather of the per-threshold int64 (tps, fps)                     -> every rank runs the same float64
     tail on the same integers as a single GPU would

The local compute steps go through a small backend object.  The product backend is ``CudaBackend``
(libmss_b200.so); the CPU test-suite injects a numpy backend built on the oracle to exercise the
host/collective logic under gloo -- there is no CPU fallback in this package.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

HIST_BITS = 16
SAMPLE_LOG2 = 24          # target number of sampled keys per rank for the splitter histogram


# peer-mapped receive buffers are expensive to set up (allocation + handle exchange): one per (device, group),
# grown collectively when a dataset needs more (the needed capacity is computed from all-gathered counts, so
# every rank takes the same decision)
_PEER_CACHE: dict = {}


class CudaBackend:
    """Local compute steps on the current CUDA device through the C ABI."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise L.MssError("CudaBackend needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        L.load()

    # -- buffers
    def new_buffer(self, capacity: int):
        from .metric import PairBuffer
        return PairBuffer(capacity, self.device)

    def append(self, buf, scores, labels, id_in, id_out):
        from .metric import _as_labels, _as_scores
        buf.append(_as_scores(scores, self.device), _as_labels(labels, self.device), id_in, id_out)

    def state(self, buf) -> Tuple[int, int, int, int]:
        return buf.read_state()

    def streams(self, buf, m, n_pos):
        """-> (negative keys, positive keys) of the buffer."""
        return buf.streams(m, n_pos)

    def finish_local(self, buf, recall_level):
        """Single-GPU result straight from the buffer (sort + counts + tail)."""
        from .metric import _finish
        return _finish(buf, recall_level)

    def _spl(self, splitters):
        return torch.from_numpy(np.asarray(list(splitters) + [0], dtype=np.uint32).view(np.int32)).to(self.device)

    # -- integer stages
    def histogram(self, keys: torch.Tensor, m: int, bits: int, every: int = 1) -> torch.Tensor:
        """Histogram of the top ``bits`` key bits over every ``every``-th group of 4 keys (1 = all keys)."""
        hist = torch.empty(1 << bits, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.load().mss_keys_histogram_sampled(keys.data_ptr(), m, bits, int(every), hist.data_ptr(),
                                                        L.stream_ptr(self.device)), "mss_keys_histogram_sampled")
        return hist

    def histogram_refine(self, keys: torch.Tensor, m: int, prefix16: int, every: int = 1) -> torch.Tensor:
        """Second level: histogram of the LOW 16 key bits of the keys whose top 16 bits equal ``prefix16`` (same sampling)."""
        hist = torch.empty(1 << 16, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.load().mss_keys_histogram_refine(keys.data_ptr(), m, int(prefix16), int(every), hist.data_ptr(),
                                                       L.stream_ptr(self.device)), "mss_keys_histogram_refine")
        return hist

    def partition(self, keys, m: int, splitters: Sequence[int], parts: int):
        import ctypes as C
        lib = L.load()
        keys_out = torch.empty(max(m, 1), dtype=torch.int32, device=self.device)
        spl = self._spl(splitters)
        counts = (C.c_int64 * parts)()
        nbytes = lib.mss_partition_workspace_bytes(m, parts)
        ws = L.workspace(nbytes, self.device)
        with torch.cuda.device(self.device):
            L.check(lib.mss_partition_keys(keys.data_ptr(), m, spl.data_ptr(), parts, keys_out.data_ptr(), counts,
                                           ws.data_ptr(), nbytes, L.stream_ptr(self.device)), "mss_partition_keys")
        return keys_out[:m], [int(c) for c in counts]

    def partition_count(self, keys, m: int, splitters: Sequence[int], parts: int) -> List[int]:
        import ctypes as C
        lib = L.load()
        spl = self._spl(splitters)
        counts = (C.c_int64 * parts)()
        ws = L.workspace(4096, self.device)
        with torch.cuda.device(self.device):
            L.check(lib.mss_partition_count(keys.data_ptr(), m, spl.data_ptr(), parts, counts, ws.data_ptr(), 4096,
                                            L.stream_ptr(self.device)), "mss_partition_count")
        return [int(c) for c in counts]

    def partition_scatter(self, keys, m: int, splitters: Sequence[int], parts: int, dst_keys: Sequence[int],
                          dst_offsets: Sequence[int]):
        """Fused partition + exchange: bucket d is stored at element offset ``dst_offsets[d]`` of the uint32 buffer at
        device address ``dst_keys[d]`` (peer memory for d != this rank)."""
        import ctypes as C
        lib = L.load()
        spl = self._spl(splitters)
        dk = (C.c_uint64 * parts)(*[int(x) for x in dst_keys])
        do = (C.c_int64 * parts)(*[int(x) for x in dst_offsets])
        nbytes = lib.mss_partition_workspace_bytes(m, parts)
        ws = L.workspace(nbytes, self.device)
        with torch.cuda.device(self.device):
            L.check(lib.mss_partition_scatter_keys(keys.data_ptr(), m, spl.data_ptr(), parts, dk, do, ws.data_ptr(), nbytes,
                                                   L.stream_ptr(self.device)), "mss_partition_scatter_keys")

    # -- both streams of a buffer per call (one host synchronisation per exchange step)
    def partition_count2(self, buf, n_neg: int, n_pos: int, splitters: Sequence[int], parts: int):
        """-> (negatives per destination, positives per destination)."""
        import ctypes as C
        lib = L.load()
        spl = (C.c_uint32 * max(parts - 1, 1))(*[int(x) for x in splitters])
        counts = (C.c_int64 * (2 * parts))()
        ws = L.workspace(16384, self.device)
        with torch.cuda.device(self.device):
            L.check(lib.mss_eval_partition_count(C.byref(buf.c), n_neg, n_pos, spl, parts, counts, ws.data_ptr(), 16384,
                                                 L.stream_ptr(self.device)), "mss_eval_partition_count")
        return [int(c) for c in counts[:parts]], [int(c) for c in counts[parts:]]

    def partition_scatter2(self, buf, n_neg: int, n_pos: int, splitters: Sequence[int], parts: int,
                           dst_keys: Sequence[int], off_neg: Sequence[int], off_pos: Sequence[int]):
        """Fused partition + exchange of both streams: destination d's buffer receives this rank's in-distribution bucket
        at element offset ``off_neg[d]`` and its OOD bucket at ``off_pos[d]`` (peer memory for d != this rank)."""
        import ctypes as C
        lib = L.load()
        spl = (C.c_uint32 * max(parts - 1, 1))(*[int(x) for x in splitters])
        dk = (C.c_uint64 * parts)(*[int(x) for x in dst_keys])
        on = (C.c_int64 * parts)(*[int(x) for x in off_neg])
        op = (C.c_int64 * parts)(*[int(x) for x in off_pos])
        nbytes = lib.mss_eval_partition_workspace_bytes(n_neg, n_pos, parts)
        ws = L.workspace(nbytes, self.device)
        with torch.cuda.device(self.device):
            L.check(lib.mss_eval_partition_scatter(C.byref(buf.c), n_neg, n_pos, spl, parts, dk, on, op, ws.data_ptr(), nbytes,
                                                   L.stream_ptr(self.device)), "mss_eval_partition_scatter")

    # -- peer-mapped receive buffers (torch symmetric memory: every rank can store into every rank's buffer)
    def peer_alloc(self, capacity: int):
        """Local half (may raise, e.g. out of memory): the symmetric allocation [keys | state], not yet mapped by the
        peers.  The receive buffer is laid out like any evaluator buffer so that peers can APPEND to it."""
        import torch.distributed._symmetric_memory as symm
        kb = (int(capacity) * 4 + 255) // 256 * 256
        return symm.empty(kb + 256, dtype=torch.uint8, device=self.device)

    def peer_map(self, raw, capacity: int, group):
        """Collective half: exchange the handles.  -> dict(buf = PairBuffer over the local part, key_ptrs / state_ptrs =
        every rank's addresses as mapped into this process, ...)."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from .metric import PairBuffer
        grp = group if group is not None else dist.group.WORLD
        if tuple(int(x) for x in torch.__version__.split("+")[0].split(".")[:2]) < (2, 8):   # older torch: enable the group first
            try:
                symm.enable_symm_mem_for_group(grp.group_name)
            except Exception:
                pass
        hdl = symm.rendezvous(raw, grp)
        cap = int(capacity)
        kb = (cap * 4 + 255) // 256 * 256
        keys = raw[: cap * 4].view(torch.int32)
        state = raw[kb: kb + L.EVAL_STATE_BYTES]
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        return {"cap": cap, "raw": raw, "hdl": hdl, "keys": keys, "buf": PairBuffer.from_tensors(keys, state, cap),
                "key_ptrs": ptrs, "state_ptrs": [p + kb for p in ptrs]}

    def exchange_append(self, buf, n_neg: int, n_pos: int, splitters: Sequence[int], parts: int, key_ptrs: Sequence[int],
                        state_ptrs: Sequence[int], capacity: int):
        """Append every key of ``buf`` to the (peer-mapped) evaluator buffer of the rank owning its key range."""
        import ctypes as C
        lib = L.load()
        spl = (C.c_uint32 * max(parts - 1, 1))(*[int(x) for x in splitters])
        dk = (C.c_uint64 * parts)(*[int(x) for x in key_ptrs])
        ds = (C.c_uint64 * parts)(*[int(x) for x in state_ptrs])
        ws = L.workspace(4096, self.device)
        with torch.cuda.device(self.device):
            L.check(lib.mss_eval_exchange_append(C.byref(buf.c), n_neg, n_pos, spl, parts, dk, ds, int(capacity), ws.data_ptr(),
                                                 4096, L.stream_ptr(self.device)), "mss_eval_exchange_append")

    def exchange_stage(self, buf, spl_dev: torch.Tensor, parts: int, out_key_ptrs: Sequence[int], out_state_ptrs: Sequence[int],
                       out_capacity: int, recv_state_ptrs: Sequence[int], recv_capacity: int, accum_state: torch.Tensor,
                       plan: torch.Tensor):
        """Enqueue (no host sync, current stream): partition a staging buffer into LOCAL outboxes, reserve the runs in the
        owners' receive states, write the copy plan to ``plan`` (pinned int64[4 * parts]), fold the staging state."""
        import ctypes as C
        lib = L.load()
        ok = (C.c_uint64 * parts)(*[int(x) for x in out_key_ptrs])
        os_ = (C.c_uint64 * parts)(*[int(x) for x in out_state_ptrs])
        rs = (C.c_uint64 * parts)(*[int(x) for x in recv_state_ptrs])
        with torch.cuda.device(self.device):
            L.check(lib.mss_eval_exchange_stage(C.byref(buf.c), spl_dev.data_ptr(), parts, ok, os_, int(out_capacity), rs,
                                                int(recv_capacity), accum_state.data_ptr(), plan.data_ptr(),
                                                L.stream_ptr(self.device)), "mss_eval_exchange_stage")

    def memcpy_async(self, dst_ptr: int, src_ptr: int, nbytes: int, stream: torch.cuda.Stream):
        L.check(L.load().mss_memcpy_async(int(dst_ptr), int(src_ptr), int(nbytes), stream.cuda_stream), "mss_memcpy_async")

    def exchange_stream(self, buf, spl_dev: torch.Tensor, parts: int, key_ptrs: Sequence[int], state_ptrs: Sequence[int],
                        capacity: int, accum_state: torch.Tensor, ws: Optional[torch.Tensor] = None):
        """Enqueue (no host sync, current stream) the remote append of a staging buffer -- sizes are read from its device
        state -- and fold its state into ``accum_state``."""
        import ctypes as C
        lib = L.load()
        dk = (C.c_uint64 * parts)(*[int(x) for x in key_ptrs])
        ds = (C.c_uint64 * parts)(*[int(x) for x in state_ptrs])
        with torch.cuda.device(self.device):
            L.check(lib.mss_eval_exchange_stream(C.byref(buf.c), spl_dev.data_ptr(), parts, dk, ds, int(capacity),
                                                 accum_state.data_ptr(), L.ptr(ws), 0 if ws is None else ws.numel(),
                                                 L.stream_ptr(self.device)), "mss_eval_exchange_stream")

    def exchange_stream_workspace(self, stage_capacity: int, parts: int) -> torch.Tensor:
        return L.workspace(L.load().mss_eval_exchange_stream_workspace_bytes(int(stage_capacity), int(parts)), self.device)

    def sort2(self, neg, n_neg: int, pos, n_pos: int):
        from .metric import sort_keys
        sort_keys(neg, n_neg, pos, n_pos)

    def counts(self, neg, n_neg: int, pos, n_pos: int, pos_before: int, neg_before: int):
        from .metric import counts_from_sorted
        return counts_from_sorted(neg, n_neg, pos, n_pos, pos_before, neg_before)

    def tail(self, tps, fps, recall_level=0.95):
        from .metric import metrics_tail
        return metrics_tail(tps, fps, recall_level)[0]

    # -- tensors used for collectives live on this device
    def tensor(self, data, dtype):
        return torch.tensor(data, dtype=dtype, device=self.device)

    def empty(self, n, dtype):
        return torch.empty(n, dtype=dtype, device=self.device)


def choose_splitters(hist: np.ndarray, world: int, bits: int = HIST_BITS) -> List[int]:
    """Key-range splitters (uint32 key values, ascending, world-1 of them) from the GLOBAL histogram of the
    top ``bits`` key bits, balancing pair counts as evenly as bin granularity allows.  Pure integer
    arithmetic on identical inputs => identical splitters on every rank."""
    hist = np.asarray(hist, dtype=np.int64)
    total = int(hist.sum())
    cum = np.cumsum(hist)
    out: List[int] = []
    for j in range(1, world):
        target = (total * j + world - 1) // world            # ceil(total * j / world)
        b = int(np.searchsorted(cum, target, side="left")) + 1   # first bin boundary with >= target pairs before it
        b = min(max(b, 0), (1 << bits))
        key = b << (32 - bits)
        out.append(min(key, 0xFFFFFFFF))
    # ascending and de-duplicated semantics are handled by dest(key) = #{j : key >= spl[j]}
    return out


def heavy_bins(hist: np.ndarray, splitters: Sequence[int], world: int, bits: int = HIST_BITS) -> List[int]:
    """Top-``bits`` bins in which a splitter should fall but cannot at bin granularity: the bin right before a splitter's
    boundary when it holds more than half of one rank's share (saturated or narrow-range scores).  Ascending, unique."""
    hist = np.asarray(hist, dtype=np.int64)
    share = max(int(hist.sum()) // max(world, 1), 1)
    out = []
    for s in splitters:
        b = min(s >> (32 - bits), 1 << bits) - 1
        if 0 <= b < (1 << bits) and int(hist[b]) * 2 > share and b not in out:
            out.append(int(b))
    return sorted(out)


def refine_splitters(hist: np.ndarray, fine: dict, world: int, bits: int = HIST_BITS) -> List[int]:
    """Splitters with a second level inside the heavy bins: ``fine[b]`` is the GLOBAL histogram of the low 32 - bits key
    bits inside top bin ``b``.  Same integer arithmetic on identical inputs on every rank => identical splitters."""
    hist = np.asarray(hist, dtype=np.int64)
    total = int(hist.sum())
    cum = np.cumsum(hist)
    low_bits = 32 - bits
    out: List[int] = []
    for j in range(1, world):
        target = (total * j + world - 1) // world
        b = int(np.searchsorted(cum, target, side="left"))                 # the bin in which the cumulative count reaches target
        b = min(b, (1 << bits) - 1)
        if b in fine:
            before = int(cum[b] - hist[b])
            fc = np.cumsum(np.asarray(fine[b], dtype=np.int64))
            # the sampled fine histogram may not add up to hist[b] exactly: scale the remaining target into it
            inside = target - before
            tot_f = int(fc[-1])
            if tot_f > 0:
                want = (inside * tot_f + int(hist[b]) - 1) // max(int(hist[b]), 1)
                f = int(np.searchsorted(fc, max(want, 1), side="left")) + 1
                key = (b << low_bits) + min(f, 1 << low_bits)
            else:
                key = (b + 1) << low_bits
        else:
            key = (b + 1) << low_bits
        out.append(min(key, 0xFFFFFFFF))
    # keep them ascending (a refined splitter can never pass the next bin boundary, but two may coincide)
    for j in range(1, len(out)):
        out[j] = max(out[j], out[j - 1])
    return out


def range_estimates(hist: np.ndarray, fine: dict, splitters: Sequence[int], bits: int = HIST_BITS) -> List[int]:
    """Expected (sampled) number of keys in each splitter range, from the two-level histogram; an upper estimate where a
    splitter falls inside a bin that has no second level."""
    hist = np.asarray(hist, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(hist)])
    low_bits = 32 - bits

    def below(key: int) -> int:                                             # keys < key
        if key >= 1 << 32:
            return int(cum[-1])
        b, low = key >> low_bits, key & ((1 << low_bits) - 1)
        if low == 0:
            return int(cum[b])
        if b in fine:
            f = np.asarray(fine[b], dtype=np.int64)
            tot = int(f.sum())
            return int(cum[b]) + (int(f[:low].sum()) * int(hist[b]) + max(tot, 1) - 1) // max(tot, 1)
        return int(cum[b + 1])

    edges = [0] + [below(int(s)) for s in splitters] + [int(cum[-1])]
    return [max(b - a, 0) for a, b in zip(edges[:-1], edges[1:])]


class StreamingEvaluator:
    """Accumulate (score, label) batches on the device; compute exact AUROC / AP / FPR@95 at the end.

    ``group``: a torch.distributed process group (or None for the default group when
    torch.distributed is initialised; single-process otherwise)."""

    def __init__(self, capacity: int, device=None, train_id_in: int = 0, train_id_out: int = 1, backend=None,
                 distributed: Optional[bool] = None, group=None, exchange: str = "auto", stage_capacity: Optional[int] = None):
        """``exchange``: "stream" = every batch is partitioned by owner right behind the kernel that scored it (local
        outboxes, same stream) and its runs are moved to the owners' peer-mapped receive buffers by the COPY ENGINES while
        the next batch is scored -- ``compute`` then starts with the keys already at their owners.  The key ranges are fixed
        from the FIRST batch (a collective inside the first ``update*`` call: every rank must make one), ``reset`` is
        collective, and a rank receives about ``capacity`` keys (+25 %).  "stream_sm" = the same with the SMs doing the
        transfer (remote append kernel on a side stream; measured slower, kept for comparison).
        "p2p" = remote append into the owners' peer-mapped buffers over NVLink (no counting pass),
        "p2p_counted" = count + all-gather of bucket sizes + fused partition/peer stores at known offsets, "nccl" = local
        partition + all-to-all, "auto" = p2p when the backend offers peer buffers (falls back if mapping fails)."""
        assert exchange in ("auto", "p2p", "p2p_counted", "nccl", "stream", "stream_sm")
        self.exchange = exchange
        self.capacity = int(capacity)
        self.stage_capacity = stage_capacity
        self._st = None                  # state of the streaming exchange (exchange="stream")
        self.backend = backend if backend is not None else CudaBackend(device)
        self.id_in, self.id_out = int(train_id_in), int(train_id_out)
        self.group = group
        if distributed is None:
            distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.distributed = bool(distributed)
        # (a streaming evaluator keeps nothing locally: batches go through two small staging buffers to their owners)
        self.buf = self.backend.new_buffer(1 if self._streaming() else int(capacity))
        self.buf.reset()

    _tags = 0

    @staticmethod
    def _next_tag():
        # evaluators are created in the same order on every rank, so the n-th streaming evaluator has tag n everywhere
        StreamingEvaluator._tags += 1
        return "stream%d" % StreamingEvaluator._tags

    # ------------------------------------------------------------------ accumulation
    def reset(self):
        """Empty the evaluator (collective in "stream" mode: every rank's receive buffer is emptied behind a barrier)."""
        self.buf.reset()
        if self._st is not None:
            self._stream_reset()

    def update(self, scores, labels):
        """One batch of score / label maps (any shape, same number of elements)."""
        if self._streaming():
            n = int(np.prod(np.shape(scores))) if not isinstance(scores, torch.Tensor) else scores.numel()
            b, buf = self._stage(n)
            self.backend.append(buf, scores, labels, self.id_in, self.id_out)
            self._stream_push(b)
            return
        self.backend.append(self.buf, scores, labels, self.id_in, self.id_out)

    def update_from_logits(self, logits: torch.Tensor, labels: torch.Tensor, key: str = "energy",
                           which: Sequence[str] = ("energy",)):
        """DeepLab fused path: score maps + ignore masking + key build in ONE kernel (deeplab.score_maps)."""
        from .deeplab import score_maps
        if self._streaming():
            b, buf = self._stage(labels.numel())
            out = score_maps(logits, which, labels=labels, evaluator=buf, key=key, id_in=self.id_in, id_out=self.id_out)
            self._stream_push(b)
            return out
        return score_maps(logits, which, labels=labels, evaluator=self.buf, key=key, id_in=self.id_in,
                          id_out=self.id_out)

    # ------------------------------------------------------------------ streaming exchange
    def _streaming(self) -> bool:
        return self.exchange in ("stream", "stream_sm") and self.distributed

    def _stage(self, numel: int):
        """-> (index, staging buffer) that the current stream may append ``numel`` pixels into."""
        be = self.backend
        if self._st is None:
            cap = int(self.stage_capacity or numel)
            dev = be.device
            self._st = {"staging": [be.new_buffer(cap), be.new_buffer(cap)], "cap": cap, "n": 0, "calibrated": False,
                        "accum": torch.zeros(L.EVAL_STATE_BYTES, dtype=torch.uint8, device=dev),
                        "side": torch.cuda.Stream(device=dev, priority=int(os.environ.get("MSS_STREAM_PRIORITY", "-1"))),
                        "appended": [torch.cuda.Event(), torch.cuda.Event()], "done": [None, None]}
            for sb in self._st["staging"]:
                sb.reset()
        st = self._st
        if numel > st["cap"]:
            raise L.MssError(f"batch of {numel} pixels exceeds the staging capacity {st['cap']} (pass stage_capacity=)")
        b = st["n"] & 1
        st["n"] += 1
        if st["done"][b] is not None:                                     # the exchange that last read this buffer
            torch.cuda.current_stream(be.device).wait_event(st["done"][b])
        return b, st["staging"][b]

    def _pick_splitters(self, hist, neg, n_neg, pos, n_pos, every, world):
        """Collective.  ``hist``: the all-reduced top-16-bit histogram.  -> (splitters, hist as numpy, {heavy bin: its
        global low-16-bit histogram}).  A splitter that would have to fall INSIDE a heavy bin (saturated or narrow-range
        scores) gets a second-level histogram of that bin, so that one rank does not end up owning -- and buffering -- the
        whole dataset; ties of one exact score still go to one rank, as the algorithm needs."""
        import torch.distributed as dist
        be, g = self.backend, self.group
        hist_np = hist.cpu().numpy()
        splitters = choose_splitters(hist_np, world)
        heavy = heavy_bins(hist_np, splitters, world) if hasattr(be, "histogram_refine") else []
        fine_by_bin: dict = {}
        if heavy:
            fine = be.empty(len(heavy) << 16, torch.int64)
            for i, b in enumerate(heavy):
                fb = be.histogram_refine(neg, n_neg, b, every)
                if n_pos:
                    fb = fb + be.histogram_refine(pos, n_pos, b, every)
                fine[i << 16: (i + 1) << 16] = fb
            dist.all_reduce(fine, group=g)
            fine_np = fine.cpu().numpy()
            fine_by_bin = {b: fine_np[i << 16: (i + 1) << 16] for i, b in enumerate(heavy)}
            splitters = refine_splitters(hist_np, fine_by_bin, world)
        return splitters, hist_np, fine_by_bin

    def _stream_calibrate(self, buf):
        """Collective, once: key ranges from the first batch, receive buffers, everything emptied behind a barrier."""
        import torch.distributed as dist
        be, g, st = self.backend, self.group, self._st
        world = dist.get_world_size(g)
        m, n_pos, _, _ = be.state(buf)
        tot = be.tensor([m, self.capacity], torch.int64)
        dist.all_reduce(tot, group=g)
        M, total_cap = (int(v) for v in tot.tolist())
        neg, pos = be.streams(buf, m, n_pos)
        every = max(1, M // (world << SAMPLE_LOG2))
        hist = be.histogram(neg, m - n_pos, HIST_BITS, every)
        if n_pos:
            hist = hist + be.histogram(pos, n_pos, HIST_BITS, every)
        dist.all_reduce(hist, group=g)
        splitters, _, _ = self._pick_splitters(hist, neg, m - n_pos, pos, n_pos, every, world)
        need = total_cap // world + total_cap // (4 * world) + (1 << 20)
        self._stream_tag = getattr(self, "_stream_tag", None) or StreamingEvaluator._next_tag()
        pb = self._peer_buffers(need, g, tag=self._stream_tag)
        if pb is None:
            raise L.MssError('exchange="stream" needs peer-mapped memory: ' + str(getattr(self, "p2p_error", "")))
        st["pb"], st["splitters"] = pb, splitters
        st["spl_dev"] = torch.from_numpy(np.asarray(list(splitters) + [0], dtype=np.uint32).view(np.int32)).to(be.device)
        pb["buf"].reset()
        torch.cuda.synchronize(be.device)
        dist.barrier(group=g)
        st["calibrated"] = True

    def _ce_push(self, b: int):
        """exchange="stream": stage batch ``b`` (partition into local outboxes + reservations + copy plan) on the current
        stream, then hand the PREVIOUS batch's runs to the copy engines -- the host waits for that batch's plan while the
        GPU is busy with this one."""
        import torch.distributed as dist
        be, st = self.backend, self._st
        world = dist.get_world_size(self.group)
        pb = st["pb"]
        if "out_keys" not in st:
            cap_o = st["cap_o"] = (st["cap"] + 3) // 4 * 4                # worst case: a whole batch for one owner (16-byte rows)
            st["out_keys"] = [be.empty(world * cap_o, torch.int32).view(world, cap_o) for _ in range(2)]
            st["out_state"] = [torch.zeros((world, L.EVAL_STATE_BYTES), dtype=torch.uint8, device=be.device) for _ in range(2)]
            st["plan"] = [torch.zeros(4 * world, dtype=torch.int64).pin_memory() for _ in range(2)]
            st["planned"], st["copied"], st["pending"], st["n_ce"] = [None, None], [None, None], None, 0
            # a few copy streams, so that the runs of a batch travel on several copy engines at once
            st["sides"] = [st["side"]] + [torch.cuda.Stream(device=be.device) for _ in range(min(3, max(world - 1, 0)))]
        p = st["n_ce"] & 1
        st["n_ce"] += 1
        cur = torch.cuda.current_stream(be.device)
        for ev in st["copied"][p] or ():
            cur.wait_event(ev)                                            # the copies that last read outbox p
        cap_o = st["cap_o"]
        ok = [st["out_keys"][p][d].data_ptr() for d in range(world)]
        os_ = [st["out_state"][p][d].data_ptr() for d in range(world)]
        be.exchange_stage(st["staging"][b], st["spl_dev"], world, ok, os_, cap_o, pb["state_ptrs"], pb["cap"], st["accum"],
                          st["plan"][p])
        ev = torch.cuda.Event()
        ev.record(cur)
        st["planned"][p] = ev
        if st["pending"] is not None:
            self._ce_flush(st["pending"])
        st["pending"] = p

    def _ce_flush(self, p: int):
        import torch.distributed as dist
        be, st = self.backend, self._st
        world = dist.get_world_size(self.group)
        pb, cap_o, sides = st["pb"], st["cap_o"], st["sides"]
        rank = dist.get_rank(self.group)
        st["planned"][p].synchronize()
        plan = st["plan"][p].tolist()
        for i in range(world):
            d = (rank + 1 + i) % world                                   # every rank starts with another owner; its own run last
            side = sides[i % len(sides)]
            cn, on, cp, op = plan[4 * d: 4 * d + 4]
            src = st["out_keys"][p][d].data_ptr()
            if cn:
                be.memcpy_async(pb["key_ptrs"][d] + 4 * on, src, 4 * cn, side)
            if cp:
                be.memcpy_async(pb["key_ptrs"][d] + 4 * op, src + 4 * (cap_o - cp), 4 * cp, side)
        evs = []
        for side in sides:
            ev = torch.cuda.Event()
            ev.record(side)
            evs.append(ev)
        st["copied"][p] = evs
        st["pending"] = None

    def _stream_push(self, b: int):
        import torch.distributed as dist
        be, st = self.backend, self._st
        if not st["calibrated"]:
            self._stream_calibrate(st["staging"][b])
        if self.exchange == "stream":
            return self._ce_push(b)
        cur = torch.cuda.current_stream(be.device)
        st["appended"][b].record(cur)
        st["side"].wait_event(st["appended"][b])
        world = dist.get_world_size(self.group)
        pb = st["pb"]
        if "ws" not in st:
            # default: every tile reserves its runs itself.  MSS_STREAM_COUNTED=1 selects the counted form (count the batch,
            # ONE remote reservation per destination and stream, look-back scatter) -- measured slower on B200: 65.5k vs
            # 69.9k images/s at 8 GPUs, 17.9k vs 20.6k at 2 (its second pass over the keys costs the co-running scoring
            # kernel more SM time than the per-tile remote atomics do)
            import os
            counted = os.environ.get("MSS_STREAM_COUNTED", "0") == "1" and world <= 16
            st["ws"] = be.exchange_stream_workspace(st["cap"], world) if counted else None
        with torch.cuda.stream(st["side"]):
            be.exchange_stream(st["staging"][b], st["spl_dev"], world, pb["key_ptrs"], pb["state_ptrs"], pb["cap"], st["accum"],
                               st["ws"])
            ev = torch.cuda.Event()
            ev.record(st["side"])
        st["done"][b] = ev

    def _stream_reset(self):
        import torch.distributed as dist
        be, st = self.backend, self._st
        if st.get("pending") is not None:
            self._ce_flush(st["pending"])
        for sd in st.get("sides", [st["side"]]):
            sd.synchronize()
        st["n"] = 0
        st["done"] = [None, None]
        if "planned" in st:
            st["planned"], st["copied"], st["pending"], st["n_ce"] = [None, None], [None, None], None, 0
        st["accum"].zero_()
        for sb in st["staging"]:
            sb.reset()
        if st["calibrated"]:
            st["pb"]["buf"].reset()
            torch.cuda.synchronize(be.device)
            dist.barrier(group=self.group)                                # nobody appends into a buffer that is being emptied

    # ------------------------------------------------------------------ result
    def compute(self, recall_level: float = 0.95):
        be = self.backend
        if not self.distributed:
            return be.finish_local(self.buf, recall_level)
        if self._streaming():
            return self._compute_streamed(recall_level)
        m, n_pos, nan, inf = be.state(self.buf)
        return self._compute_distributed(m, n_pos, nan, inf, recall_level)

    def _compute_streamed(self, recall_level):
        """exchange="stream": the keys are already at their owners; what is left is sort + counts + tail."""
        import time
        import torch.distributed as dist
        be, g, st = self.backend, self.group, self._st
        world, rank = dist.get_world_size(g), dist.get_rank(g)
        marks = [("start", time.perf_counter())]

        def mark(name):
            torch.cuda.synchronize(be.device)
            marks.append((name, time.perf_counter()))

        if st is None or not st["calibrated"]:
            raise L.MssError('exchange="stream": compute() before any update (every rank must feed at least one batch)')
        if st.get("pending") is not None:
            self._ce_flush(st["pending"])                                 # the last batch's runs
        for sd in st.get("sides", [st["side"]]):
            sd.synchronize()
        torch.cuda.synchronize(be.device)
        dist.barrier(group=g)                                             # every rank's appends have landed
        acc = st["accum"].cpu().numpy()
        n_neg, n_pos = (int(v) for v in acc[:16].view(np.uint64))
        nan, inf = (int(v) for v in acc[16:24].view(np.uint32))
        dropped = int(acc[24:32].view(np.uint64)[0])
        try:
            m2, m2_pos, _, _ = st["pb"]["buf"].read_state()
            over = 0
        except L.MssError:
            m2 = m2_pos = 0
            over = 1
        mine = be.tensor([n_neg + n_pos, n_pos, nan, inf, m2 - m2_pos, m2_pos, over + (1 if dropped else 0)], torch.int64)
        got = be.empty(7 * world, torch.int64)
        dist.all_gather_into_tensor(got, mine, group=g)
        got = got.view(world, 7).cpu().numpy()
        M, P, gnan, ginf = (int(v) for v in got[:, :4].sum(axis=0))
        if int(got[:, 6].max()):
            raise L.MssError('exchange="stream": a staging or receive buffer overflowed (the key ranges fixed from the first '
                             'batch did not balance this dataset): raise capacity / stage_capacity, or use exchange="p2p"')
        if P == 0 or P == M:
            return None
        _raise_nonfinite(gnan, ginf)
        mark("barrier_and_sizes")
        per_dst = got[:, 4:6].T                                           # [stream, dst]
        r_neg, r_pos = st["pb"]["buf"].streams(m2, m2_pos)
        return self._sort_count_tail(r_neg, m2 - m2_pos, r_pos, m2_pos, per_dst, recall_level, marks, mark,
                                     {"send_counts": None, "recv_counts": [[m2 - m2_pos], [m2_pos]],
                                      "splitters": st["splitters"], "exchange": self.exchange})

    def _peer_buffers(self, need: int, g, tag=None):
        """Peer-mapped receive buffer of >= need keys on every rank, or None (on EVERY rank) if it cannot be had.
        Capability and the local allocation are agreed on with a cheap collective BEFORE the collective handle
        exchange, so a rank that fails (no peer access, out of memory) cannot leave the others blocked in it."""
        import torch.distributed as dist
        be = self.backend
        grp = g if g is not None else dist.group.WORLD
        # (a streaming evaluator keeps its receive buffer for a whole round: it gets one of its own, `tag`)
        ck = (str(getattr(be, "device", "cpu")), getattr(grp, "group_name", None) or id(grp), tag)
        cached = _PEER_CACHE.get(ck)
        # `need` comes from all-gathered counts and every rank has taken the same decisions before, so the cache state
        # is the same everywhere; the MIN all-reduce makes that an agreement instead of an assumption
        have = be.tensor([1 if (cached is not None and cached["cap"] >= need) else 0], torch.int64)
        dist.all_reduce(have, op=dist.ReduceOp.MIN, group=g)
        if int(have.item()) == 1:
            return cached
        cap = max(need + need // 8, 1 << 20)
        raw, err = None, None
        try:
            raw = be.peer_alloc(cap)
        except Exception as e:
            err = repr(e)
        ok = be.tensor([0 if raw is None else 1], torch.int64)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=g)
        if int(ok.item()) != 1:
            self.p2p_error = err or "a peer could not allocate its receive buffer"
            return None
        try:
            _PEER_CACHE[ck] = be.peer_map(raw, cap, g)
        except Exception as e:               # the handle exchange itself is collective: it fails (or not) everywhere
            self.p2p_error = repr(e)
            return None
        return _PEER_CACHE[ck]

    def _compute_distributed(self, m, n_pos, nan, inf, recall_level):
        import time
        import torch.distributed as dist
        be, g = self.backend, self.group
        world, rank = dist.get_world_size(g), dist.get_rank(g)
        # host wall clock per phase (every phase ends in a host-visible result, so it is synchronised)
        marks = [("start", time.perf_counter())]

        def mark(name):
            if torch.cuda.is_available() and getattr(be, "device", None) is not None:
                torch.cuda.synchronize(be.device)
            marks.append((name, time.perf_counter()))

        # 1. global emptiness / finiteness decisions (identical on every rank)
        tot = be.tensor([m, n_pos, nan, inf], torch.int64)
        dist.all_reduce(tot, group=g)
        M, P, gnan, ginf = (int(v) for v in tot.tolist())
        if P == 0 or P == M:
            return None
        _raise_nonfinite(gnan, ginf)
        mark("counts_allreduce")

        # 2. global histogram of the top key bits (both streams) -> splitters
        neg, pos = be.streams(self.buf, m, n_pos)
        n_neg = m - n_pos
        # a systematic sample of ~2^24 keys per rank is plenty to balance the ranges; any splitters give the exact
        # result as long as every rank uses the same ones, which the all_reduce guarantees
        every = max(1, M // (world << SAMPLE_LOG2))
        hist = be.histogram(neg, n_neg, HIST_BITS, every)
        if n_pos:
            hist = hist + be.histogram(pos, n_pos, HIST_BITS, every)
        dist.all_reduce(hist, group=g)
        splitters, hist_np, fine_by_bin = self._pick_splitters(hist, neg, n_neg, pos, n_pos, every, world)
        heavy = sorted(fine_by_bin)
        mark("histogram_splitters")

        # 3. exchange by key range.  Receive layout on rank r = the evaluator's own two-stream layout.
        exchange = self.exchange
        if exchange == "auto":
            # measured on B200 / NVSwitch (cfg-4): remote append 8.6 ms vs count + scatter 10.7 ms at 2 GPUs, but 6.5 vs
            # 5.8 ms at 8 (a remote atomic per tile and destination: 8 M of them in flight across the switch)
            p2p = "p2p" if world <= 4 else "p2p_counted"
            exchange = p2p if hasattr(be, "peer_alloc") and not getattr(self, "_p2p_failed", False) else "nccl"
        r_neg = r_pos = None
        if exchange == "p2p":
            # 3a. remote append: nothing is counted beforehand.  The receive capacity comes from the (sampled) global
            #     histogram: expected keys of the fullest key range + 8 % + 64 K.
            est = max(range_estimates(hist_np, fine_by_bin, splitters)) * every
            need = est + est // 12 + (1 << 16)
            pb = self._peer_buffers(need, g)
            if pb is None:
                self._p2p_failed = True
                exchange = "nccl"
            else:
                rb = pb["buf"]
                rb.reset()
                torch.cuda.synchronize(be.device)
                dist.barrier(group=g)                                     # every receive buffer is empty
                be.exchange_append(self.buf, n_neg, n_pos, splitters, world, pb["key_ptrs"], pb["state_ptrs"], pb["cap"])
                dist.barrier(group=g)                                     # every rank's appends have landed
                try:
                    m2, m2_pos, _, _ = rb.read_state()
                    over = 0
                except L.MssError:                                        # a receive buffer overflowed (estimate too small)
                    m2 = m2_pos = 0
                    over = 1
                mine = be.tensor([m2 - m2_pos, m2_pos, over], torch.int64)
                got = be.empty(3 * world, torch.int64)
                dist.all_gather_into_tensor(got, mine, group=g)
                got = got.view(world, 3).cpu().numpy()
                if int(got[:, 2].max()):
                    exchange = "p2p_counted"                              # agreed on by all ranks: exact sizes first
                else:
                    m2_neg = m2 - m2_pos
                    per_dst = got[:, :2].T                                # [stream, dst]
                    send, recv_neg, recv_pos = None, [m2_neg], [m2_pos]
                    r_neg, r_pos = rb.streams(m2, m2_pos)
                    mark("exchange_append_p2p")
        if exchange in ("p2p_counted", "nccl"):
            if exchange == "p2p_counted":
                send = list(be.partition_count2(self.buf, n_neg, n_pos, splitters, world))
            else:
                pk_neg, c_neg = be.partition(neg, n_neg, splitters, world)
                pk_pos, c_pos = be.partition(pos, n_pos, splitters, world)
                send = [c_neg, c_pos]
                mark("partition")
            cm = be.tensor(send[0] + send[1], torch.int64)
            all_counts = be.empty(world * 2 * world, torch.int64)
            dist.all_gather_into_tensor(all_counts, cm, group=g)
            all_counts = all_counts.view(world, 2, world).cpu().numpy()       # [src, stream, dst]
            recv_neg = [int(c) for c in all_counts[:, 0, rank]]
            recv_pos = [int(c) for c in all_counts[:, 1, rank]]
            m2_neg, m2_pos = int(sum(recv_neg)), int(sum(recv_pos))
            per_dst = all_counts.sum(axis=0)                                  # [stream, dst]
            need = int((per_dst[0] + per_dst[1]).max())                       # identical on every rank
            if exchange == "p2p_counted":
                pb = self._peer_buffers(need, g)
                if pb is None:                                                # agreed on by all ranks: NCCL path
                    self._p2p_failed = True
                    exchange = "nccl"
                    pk_neg, _ = be.partition(neg, n_neg, splitters, world)
                    pk_pos, _ = be.partition(pos, n_pos, splitters, world)
                    mark("partition")
                else:
                    mark("count")
            if exchange == "p2p_counted":
                # 3b. fused: ONE call partitions both streams and stores every bucket straight into its owner's receive
                #     buffer over NVLink at offsets known from the all-gathered counts
                cap = pb["cap"]
                off_neg = [int(all_counts[:rank, 0, d].sum()) for d in range(world)]
                off_pos = [cap - int(per_dst[1][d]) + int(all_counts[:rank, 1, d].sum()) for d in range(world)]
                torch.cuda.synchronize(be.device)
                dist.barrier(group=g)                                         # peers are done with the previous contents
                be.partition_scatter2(self.buf, n_neg, n_pos, splitters, world, pb["key_ptrs"], off_neg, off_pos)
                dist.barrier(group=g)                                         # every rank's stores have landed
                rk = pb["keys"]
                r_neg, r_pos = rk[:m2_neg], rk[cap - m2_pos: cap]
                mark("partition_scatter_p2p")
            else:
                # 3c. local partition by destination rank + NCCL all-to-all per stream
                r_neg = be.empty(max(m2_neg, 1), torch.int32)[:m2_neg]
                r_pos = be.empty(max(m2_pos, 1), torch.int32)[:m2_pos]
                dist.all_to_all_single(r_neg, pk_neg, recv_neg, send[0], group=g)
                dist.all_to_all_single(r_pos, pk_pos, recv_pos, send[1], group=g)
                mark("all_to_all")

        return self._sort_count_tail(r_neg, m2_neg, r_pos, m2_pos, per_dst, recall_level, marks, mark,
                                     {"send_counts": send, "recv_counts": [recv_neg, recv_pos], "splitters": splitters,
                                      "exchange": exchange, "refined_bins": heavy,
                                      "dst_totals": [int(a) + int(b) for a, b in zip(per_dst[0], per_dst[1])]})

    def _sort_count_tail(self, r_neg, m2_neg, r_pos, m2_pos, per_dst, recall_level, marks, mark, info):
        import torch.distributed as dist
        be, g = self.backend, self.group
        world, rank = dist.get_world_size(g), dist.get_rank(g)
        # 4. local sort of both streams + merge-path counts with the global prefixes (known from the count matrix)
        be.sort2(r_neg, m2_neg, r_pos, m2_pos)
        mark("sort")
        neg_before = int(per_dst[0][:rank].sum())
        pos_before = int(per_dst[1][:rank].sum())
        if m2_neg + m2_pos:
            tps, fps = be.counts(r_neg, m2_neg, r_pos, m2_pos, pos_before, neg_before)
        else:
            tps, fps = be.empty(0, torch.int64), be.empty(0, torch.int64)
        mark("counts")

        # 5. gather every rank's thresholds (padded to the longest slice), run the identical tail everywhere
        T = be.tensor([tps.numel()], torch.int64)
        Ts = be.empty(world, torch.int64)
        dist.all_gather_into_tensor(Ts, T, group=g)
        Ts = [int(v) for v in Ts.tolist()]
        Tmax = max(max(Ts), 1)
        pad = be.empty(2 * Tmax, torch.int64)
        pad.zero_()
        pad[: tps.numel()] = tps
        pad[Tmax: Tmax + fps.numel()] = fps
        gathered = be.empty(2 * Tmax * world, torch.int64)
        dist.all_gather_into_tensor(gathered, pad, group=g)
        gathered = gathered.view(world, 2, Tmax)
        tps_all = torch.cat([gathered[r, 0, : Ts[r]] for r in range(world)])
        fps_all = torch.cat([gathered[r, 1, : Ts[r]] for r in range(world)])
        mark("gather_thresholds")
        res = be.tail(tps_all, fps_all, recall_level)
        mark("tail")
        info.update({"thresholds_per_rank": Ts, "p2p_error": getattr(self, "p2p_error", None),
                     "phase_ms": {n: (t - marks[i][1]) * 1e3 for i, (n, t) in enumerate(marks[1:])}})
        self.last_exchange = info
        return res


def _raise_nonfinite(nan, inf):
    if nan:
        raise ValueError("Input contains NaN.")
    if inf:
        raise ValueError("Input contains infinity or a value too large for dtype('float32').")
