"""Streaming, multi-GPU evaluator: the on-device replacement of the reference's tester loop
(test_deeplab.py:84-117, test_m2f.py:125-158, and ``valid_batch`` in both trainers), i.e.

    anomaly_scores.append(score.cpu().numpy()); ood_gts.append(target.cpu().numpy())   # per batch
    eval_ood_measure(np.concatenate(anomaly_scores), np.concatenate(ood_gts))           # at the end

``update`` appends the valid pixels of a batch to a device buffer of order-preserving keys (no D2H);
``compute`` returns exactly what ``eval_ood_measure`` returns for the concatenated dataset.

Multi-GPU (one process per GPU, images sharded over ranks, ``torch.distributed``): the only exchange
step of the whole path.  Integer-exact, so the N-GPU result is bit-identical to the 1-GPU result:

  1. all_reduce of (count, positives, nan, inf)                          -> None / ValueError decisions
  2. all_reduce of a 2^16-bin histogram of the top key bits              -> identical splitters everywhere
  3. local stable partition by key range + all_to_all of (key, label)    -> rank r owns key range r
     (a distinct score never straddles two ranks)
  4. local radix sort, run-length counts with global (index, positives) prefixes from an all_gather
  5. all_gather of the per-threshold int64 (tps, fps)                     -> every rank runs the same float64
     tail on the same integers as a single GPU would

The local compute steps go through a small backend object.  The product backend is ``CudaBackend``
(libmss_b200.so); the CPU test-suite injects a numpy backend built on the oracle to exercise the
host/collective logic under gloo -- there is no CPU fallback in this package.
"""
from __future__ import annotations

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

    def pairs(self, buf, m):
        return buf.keys[:m], buf.labs[:m]

    # -- integer stages
    def histogram(self, keys: torch.Tensor, m: int, bits: int, every: int = 1) -> torch.Tensor:
        """Histogram of the top ``bits`` key bits over every ``every``-th group of 4 keys (1 = all keys)."""
        hist = torch.empty(1 << bits, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.load().mss_keys_histogram_sampled(keys.data_ptr(), m, bits, int(every), hist.data_ptr(),
                                                        L.stream_ptr(self.device)), "mss_keys_histogram_sampled")
        return hist

    def partition(self, keys, labs, m: int, splitters: Sequence[int], parts: int):
        import ctypes as C
        lib = L.load()
        keys_out = torch.empty(max(m, 1), dtype=torch.int32, device=self.device)
        labs_out = torch.empty(max(m, 1), dtype=torch.uint8, device=self.device)
        spl = torch.from_numpy(np.asarray(list(splitters) + [0], dtype=np.uint32).view(np.int32)).to(self.device)
        counts = (C.c_int64 * parts)()
        nbytes = lib.mss_partition_workspace_bytes(m)
        ws = L.workspace(nbytes, self.device)
        with torch.cuda.device(self.device):
            L.check(lib.mss_partition_pairs(keys.data_ptr(), labs.data_ptr(), m, spl.data_ptr(), parts,
                                            keys_out.data_ptr(), labs_out.data_ptr(), counts, ws.data_ptr(), nbytes,
                                            L.stream_ptr(self.device)), "mss_partition_pairs")
        return keys_out[:m], labs_out[:m], [int(c) for c in counts]

    def partition_count(self, keys, m: int, splitters: Sequence[int], parts: int) -> List[int]:
        import ctypes as C
        lib = L.load()
        spl = torch.from_numpy(np.asarray(list(splitters) + [0], dtype=np.uint32).view(np.int32)).to(self.device)
        counts = (C.c_int64 * parts)()
        ws = L.workspace(4096, self.device)
        with torch.cuda.device(self.device):
            L.check(lib.mss_partition_count(keys.data_ptr(), m, spl.data_ptr(), parts, counts, ws.data_ptr(), 4096,
                                            L.stream_ptr(self.device)), "mss_partition_count")
        return [int(c) for c in counts]

    def partition_scatter(self, keys, labs, m: int, splitters: Sequence[int], parts: int, dst_keys: Sequence[int],
                          dst_labs: Sequence[int], dst_offsets: Sequence[int]):
        """Fused partition + exchange: bucket d is stored at element offset ``dst_offsets[d]`` of the buffers at
        device addresses ``dst_keys[d]`` / ``dst_labs[d]`` (peer memory for d != this rank)."""
        import ctypes as C
        lib = L.load()
        spl = torch.from_numpy(np.asarray(list(splitters) + [0], dtype=np.uint32).view(np.int32)).to(self.device)
        dk = (C.c_uint64 * parts)(*[int(x) for x in dst_keys])
        dl = (C.c_uint64 * parts)(*[int(x) for x in dst_labs])
        do = (C.c_int64 * parts)(*[int(x) for x in dst_offsets])
        nbytes = lib.mss_partition_workspace_bytes(m)
        ws = L.workspace(nbytes, self.device)
        with torch.cuda.device(self.device):
            L.check(lib.mss_partition_scatter_pairs(keys.data_ptr(), labs.data_ptr(), m, spl.data_ptr(), parts, dk, dl, do,
                                                    ws.data_ptr(), nbytes, L.stream_ptr(self.device)),
                    "mss_partition_scatter_pairs")

    # -- peer-mapped receive buffers (torch symmetric memory: every rank can store into every rank's buffer)
    def peer_buffers(self, capacity: int, group):
        """-> (keys int32 [capacity], labs uint8 [capacity], key_ptrs[world], lab_ptrs[world], handle) or raises."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        cap = int(capacity)
        ck = (str(self.device), id(group))
        cached = _PEER_CACHE.get(ck)
        if cached is not None and cached["cap"] >= cap:
            return cached
        grp = group if group is not None else dist.group.WORLD
        if tuple(int(x) for x in torch.__version__.split("+")[0].split(".")[:2]) < (2, 8):   # older torch: enable the group first
            try:
                symm.enable_symm_mem_for_group(grp.group_name)
            except Exception:
                pass
        kb = (cap * 4 + 255) // 256 * 256
        raw = symm.empty(kb + cap, dtype=torch.uint8, device=self.device)     # one allocation: [keys | labs]
        hdl = symm.rendezvous(raw, grp)
        keys = raw[: cap * 4].view(torch.int32)
        labs = raw[kb: kb + cap]
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        _PEER_CACHE[ck] = {"cap": cap, "raw": raw, "hdl": hdl, "keys": keys, "labs": labs,
                           "key_ptrs": ptrs, "lab_ptrs": [p + kb for p in ptrs]}
        return _PEER_CACHE[ck]

    def sort(self, keys, labs, m: int):
        from .metric import sort_pairs
        sort_pairs(keys, labs, m)

    def counts(self, keys, labs, m: int, pos_before: int, idx_before: int):
        from .metric import counts_from_sorted
        tps, fps, _, _ = counts_from_sorted(keys, labs, m, pos_before, idx_before)
        return tps, fps

    def counts_local(self, keys, labs, m: int):
        """-> (tps, fps, #positives) of this slice alone (prefixes of earlier slices not included)."""
        from .metric import counts_from_sorted
        tps, fps, n_pos, _ = counts_from_sorted(keys, labs, m, 0, 0)
        return tps, fps, n_pos

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


class StreamingEvaluator:
    """Accumulate (score, label) batches on the device; compute exact AUROC / AP / FPR@95 at the end.

    ``group``: a torch.distributed process group (or None for the default group when
    torch.distributed is initialised; single-process otherwise)."""

    def __init__(self, capacity: int, device=None, train_id_in: int = 0, train_id_out: int = 1, backend=None,
                 distributed: Optional[bool] = None, group=None, exchange: str = "auto"):
        """``exchange``: "p2p" = fused partition + peer-memory stores over NVLink, "nccl" = local partition +
        all-to-all, "auto" = p2p when the backend offers peer buffers (falls back to nccl if mapping fails)."""
        assert exchange in ("auto", "p2p", "nccl")
        self.exchange = exchange
        self.backend = backend if backend is not None else CudaBackend(device)
        self.id_in, self.id_out = int(train_id_in), int(train_id_out)
        self.buf = self.backend.new_buffer(int(capacity))
        self.buf.reset()
        self.group = group
        if distributed is None:
            distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.distributed = bool(distributed)

    # ------------------------------------------------------------------ accumulation
    def reset(self):
        self.buf.reset()

    def update(self, scores, labels):
        """One batch of score / label maps (any shape, same number of elements)."""
        self.backend.append(self.buf, scores, labels, self.id_in, self.id_out)

    def update_from_logits(self, logits: torch.Tensor, labels: torch.Tensor, key: str = "energy",
                           which: Sequence[str] = ("energy",)):
        """DeepLab fused path: score maps + ignore masking + key build in ONE kernel (deeplab.score_maps)."""
        from .deeplab import score_maps
        return score_maps(logits, which, labels=labels, evaluator=self.buf, key=key, id_in=self.id_in,
                          id_out=self.id_out)

    # ------------------------------------------------------------------ result
    def compute(self, recall_level: float = 0.95):
        be = self.backend
        m, n_pos, nan, inf = be.state(self.buf)
        if not self.distributed:
            if n_pos == 0 or n_pos == m:
                return None
            _raise_nonfinite(nan, inf)
            keys, labs = be.pairs(self.buf, m)
            be.sort(keys, labs, m)
            tps, fps = be.counts(keys, labs, m, 0, 0)
            return be.tail(tps, fps, recall_level)
        return self._compute_distributed(m, n_pos, nan, inf, recall_level)

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

        # 2. global histogram of the top key bits -> splitters
        keys, labs = be.pairs(self.buf, m)
        # a systematic sample of ~2^24 keys per rank is plenty to balance the ranges; any splitters give the exact
        # result as long as every rank uses the same ones, which the all_reduce guarantees
        every = max(1, M // (world << SAMPLE_LOG2))
        hist = be.histogram(keys, m, HIST_BITS, every)
        dist.all_reduce(hist, group=g)
        splitters = choose_splitters(hist.cpu().numpy(), world)
        mark("histogram_splitters")

        # 3. exchange by key range
        exchange = self.exchange
        if exchange == "auto":
            exchange = "p2p" if hasattr(be, "peer_buffers") and not getattr(self, "_p2p_failed", False) else "nccl"
        if exchange == "p2p":
            # 3a. fused: count -> all-gather of the counts -> ONE kernel that partitions and stores every bucket
            #     straight into its owner's receive buffer over NVLink (no staging copy, no all-to-all)
            send_counts = be.partition_count(keys, m, splitters, world)
            cm = be.tensor(send_counts, torch.int64)
            all_counts = be.empty(world * world, torch.int64)
            dist.all_gather_into_tensor(all_counts, cm, group=g)
            all_counts = all_counts.view(world, world).cpu().numpy()      # [src, dst]
            recv_counts = [int(c) for c in all_counts[:, rank]]
            m2 = int(sum(recv_counts))
            need = int(all_counts.sum(axis=0).max())                      # identical on every rank
            try:
                pb = be.peer_buffers(max(need + need // 8, 1 << 20), g)   # collective (re)allocation when it grows
            except Exception as e:                                        # no peer access on this box: NCCL path
                self._p2p_failed, self.p2p_error = True, repr(e)
                pb = None
            ok = be.tensor([0 if pb is None else 1], torch.int64)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=g)
            if int(ok.item()) == 1:
                mark("count")
                offsets = [int(all_counts[:rank, d].sum()) for d in range(world)]   # my block inside rank d's buffer
                dist.barrier(group=g)                                     # peers are done with the previous contents
                be.partition_scatter(keys, labs, m, splitters, world, pb["key_ptrs"], pb["lab_ptrs"], offsets)
                dist.barrier(group=g)                                     # every rank's stores have landed
                rk, rl = pb["keys"], pb["labs"]
                mark("partition_scatter_p2p")
            else:
                self._p2p_failed = True
                exchange = "nccl"
        if exchange == "nccl":
            # 3b. local partition by destination rank + NCCL all-to-all
            pk, pl, send_counts = be.partition(keys, labs, m, splitters, world)
            mark("partition")
            cm = be.tensor(send_counts, torch.int64)
            all_counts = be.empty(world * world, torch.int64)
            dist.all_gather_into_tensor(all_counts, cm, group=g)
            all_counts = all_counts.view(world, world).cpu().numpy()      # [src, dst]
            recv_counts = [int(c) for c in all_counts[:, rank]]
            m2 = int(sum(recv_counts))
            rk, rl = be.empty(max(m2, 1), torch.int32), be.empty(max(m2, 1), torch.uint8)
            dist.all_to_all_single(rk[:m2], pk, recv_counts, send_counts, group=g)
            dist.all_to_all_single(rl[:m2], pl, recv_counts, send_counts, group=g)
            mark("all_to_all")

        # 4. local sort + run-length counts with global prefixes
        be.sort(rk, rl, m2)
        mark("sort")
        # local cumulative counts first (they also yield this slice's #positives), global prefixes added afterwards:
        # tps += positives before this rank, fps += negatives before this rank
        if m2:
            tps, fps, lp = be.counts_local(rk, rl, m2)
        else:
            tps, fps, lp = be.empty(0, torch.int64), be.empty(0, torch.int64), 0
        mine = be.tensor([m2, lp], torch.int64)
        per_rank = be.empty(2 * world, torch.int64)
        dist.all_gather_into_tensor(per_rank, mine, group=g)
        per_rank = per_rank.view(world, 2).cpu().numpy()
        idx_before = int(per_rank[:rank, 0].sum())
        pos_before = int(per_rank[:rank, 1].sum())
        if tps.numel():
            tps += pos_before
            fps += idx_before - pos_before
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
        self.last_exchange = {"send_counts": send_counts, "recv_counts": recv_counts, "splitters": splitters,
                              "thresholds_per_rank": Ts, "exchange": exchange,
                              "p2p_error": getattr(self, "p2p_error", None),
                              "phase_ms": {n: (t - marks[i][1]) * 1e3 for i, (n, t) in enumerate(marks[1:])}}
        return res


def _raise_nonfinite(nan, inf):
    if nan:
        raise ValueError("Input contains NaN.")
    if inf:
        raise ValueError("Input contains infinity or a value too large for dtype('float32').")
