"""Host / collective logic of the multi-GPU evaluator under gloo (world_size 2 and 3, CPU): the local compute
steps are served by the oracle-backed NumpyBackend, everything else (splitters, all_to_all plan, prefix
bookkeeping, gather + tail) is the product code.  The distributed result must be bit-identical to the
single-process result and to the reference-pinned oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gen_inputs as gi
from multishiftseg_b200.evaluator import (StreamingEvaluator, choose_splitters, heavy_bins, range_estimates,
                                          refine_splitters)
from numpy_backend import NumpyBackend
from oracle import c_oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


CASES = [("cont", 20011, 0.1, 0.05), ("q2", 30007, 0.05, 0.2), ("const", 5000, 0.3, 0.0), ("f16", 40009, 0.01, 0.05),
         ("zeros", 9000, 0.3, 0.05)]


def _narrow_case(n=30011):
    """Scores in one top-16-bit key bin (0.999 .. 1.0) plus a saturated tie mass: the case bin-granular splitters cannot
    balance."""
    rng = np.random.default_rng(77)
    s = (0.999 + 0.001 * rng.random(n)).astype(np.float32)
    s[rng.random(n) < 0.2] = np.float32(0.9995)
    l = (rng.random(n) < 0.1).astype(np.int64)
    l[rng.random(n) < 0.05] = 255
    return s, l


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for ci, (mode, n, p_ood, p_ign) in enumerate(CASES):
        s, l = gi.metric_case(100 + ci, n, mode, p_ood, p_ign)
        ev = StreamingEvaluator(n, backend=NumpyBackend(), distributed=True)
        # images sharded round-robin in chunks ("images") of 1000 pixels, two updates per rank
        chunks = [(i, min(i + 1000, n)) for i in range(0, n, 1000)]
        mine = chunks[rank::world]
        for a, b in mine:
            ev.update(s[a:b], l[a:b])
        r = ev.compute()
        out.append(None if r is None else tuple(float(v) for v in r))
    s, l = _narrow_case()
    ev = StreamingEvaluator(len(s), backend=NumpyBackend(), distributed=True)
    ev.update(s[rank::world], l[rank::world])
    r = ev.compute()
    out.append(tuple(float(v) for v in r))
    # a rank with no data at all, and an all-ignored dataset
    ev = StreamingEvaluator(10, backend=NumpyBackend(), distributed=True)
    if rank == 0:
        ev.update(np.array([.1, .4, .35, .8], np.float32), np.array([0, 0, 1, 1]))
    r = ev.compute()
    out.append(tuple(float(v) for v in r))
    ev = StreamingEvaluator(10, backend=NumpyBackend(), distributed=True)
    ev.update(np.array([.1, .4], np.float32), np.array([255, 0]))
    out.append(ev.compute())
    # NaN on one rank raises on every rank
    ev = StreamingEvaluator(10, backend=NumpyBackend(), distributed=True)
    ev.update(np.array([.1, np.nan if rank == world - 1 else .2, .3], np.float32), np.array([0, 1, 1]))
    try:
        ev.compute()
        out.append("no error")
    except ValueError as e:
        out.append(str(e))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_equals_single_process_and_oracle(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = []
    for ci, (mode, n, p_ood, p_ign) in enumerate(CASES):
        s, l = gi.metric_case(100 + ci, n, mode, p_ood, p_ign)
        want.append(c_oracle.eval_ood_measure(s, l))
        single = StreamingEvaluator(n, backend=NumpyBackend(), distributed=False)
        single.update(s, l)
        assert tuple(float(v) for v in single.compute()) == want[-1]
    want.append(c_oracle.eval_ood_measure(*_narrow_case()))
    want.append((0.75, float.fromhex("0x1.aaaaaaaaaaaaap-1"), 0.5))     # K1
    want.append(None)
    want.append("Input contains NaN.")
    for r in range(world):
        assert results[r] == want, (r, results[r], want)


def test_choose_splitters_properties():
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        hist = rng.integers(0, 1000, size=1 << 16)
        hist[rng.integers(0, 1 << 16, size=60000)] = 0
        spl = choose_splitters(hist, world)
        assert len(spl) == world - 1 and spl == sorted(spl)
        assert all(s % (1 << 16) == 0 or s == 0xFFFFFFFF for s in spl)
        # balance: every rank gets at most its share plus one bin
        bounds = [0] + [s >> 16 for s in spl] + [1 << 16]
        loads = [int(hist[a:b].sum()) for a, b in zip(bounds[:-1], bounds[1:])]
        assert sum(loads) == int(hist.sum())
        assert max(loads) <= int(hist.sum()) / world + int(hist.max()) + 1
    # degenerate: everything in one bin -> one rank takes it all, still a valid ascending split
    hist = np.zeros(1 << 16, np.int64)
    hist[12345] = 10 ** 9
    spl = choose_splitters(hist, 4)
    assert spl == sorted(spl)


def test_refined_splitters_balance_a_single_heavy_bin():
    from oracle import metrics_oracle
    s, l = _narrow_case(200000)
    keys = metrics_oracle.float_key_desc(s[l != 255]).astype(np.uint32)
    hist = np.bincount(keys >> 16, minlength=1 << 16).astype(np.int64)
    for world in (2, 4, 8):
        coarse = choose_splitters(hist, world)
        heavy = heavy_bins(hist, coarse, world)
        assert heavy, "the narrow case must trigger the second level"
        fine = {b: np.bincount(keys[(keys >> 16) == b] & 0xFFFF, minlength=1 << 16).astype(np.int64) for b in heavy}
        spl = refine_splitters(hist, fine, world)
        assert len(spl) == world - 1 and spl == sorted(spl)
        dest = np.searchsorted(np.asarray(spl, dtype=np.uint64), keys.astype(np.uint64), side="right")
        loads = np.bincount(dest, minlength=world)
        biggest_tie = int(np.unique(keys, return_counts=True)[1].max())
        assert loads.max() <= len(keys) // world + biggest_tie + 2, (world, loads.tolist())
        est = range_estimates(hist, fine, spl)
        assert sum(est) == len(keys)
        assert all(abs(int(e) - int(c)) <= 2 for e, c in zip(est, loads)), (est, loads.tolist())
        # without the second level one rank would own (almost) everything
        d0 = np.searchsorted(np.asarray(coarse, dtype=np.uint64), keys.astype(np.uint64), side="right")
        assert np.bincount(d0, minlength=world).max() > 0.9 * len(keys)


# ---- the STREAMED exchange (what bench.py runs for N > 1) under gloo: shared memory stands in for the peer-mapped
#      buffers, a file lock for the system-scope atomics, memmove for the copy engines (tests/fake_peer_backend.py)
def _stream_worker(rank, world, port, lock_path, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import fake_peer_backend as fp
    fp.install_cuda_stubs()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    be = fp.FakePeerBackend(lock_path)
    out = []
    try:
        for ci, (mode, n, p_ood, p_ign) in enumerate(CASES[:4]):
            s, l = gi.metric_case(100 + ci, n, mode, p_ood, p_ign)
            img = 1003                                              # not a multiple of 4: outbox rows get padded
            chunks = [(i, min(i + img, n)) for i in range(0, n, img)]
            ev = StreamingEvaluator(n // world + 4 * img, backend=be, distributed=True, exchange="stream", stage_capacity=img)
            for rep in range(2):                                    # the second round reuses key ranges and buffers
                ev.reset()
                mine = chunks[rank::world]
                if rep == 1 and rank == world - 1:
                    mine = mine[::-1]                               # another arrival order, same multiset
                for a, b in mine:
                    ev.update(s[a:b], l[a:b])
                if rank == 0:                                       # a batch without a single valid pixel
                    ev.update(np.zeros(7, np.float32), np.full(7, 255))
                r = ev.compute()
                out.append(None if r is None else tuple(float(v) for v in r))
            out.append(("refined", bool(ev._st and ev._st.get("splitters"))))
        out.append(("copies", be.copies > 0))
        # a batch larger than the staging capacity is refused (before anything collective happens)
        s, l = gi.metric_case(100, 20011, "cont", 0.1, 0.05)
        ev = StreamingEvaluator(2000, backend=be, distributed=True, exchange="stream", stage_capacity=100)
        try:
            ev.update(s[:1003], l[:1003])
            out.append("no error")
        except Exception as e:
            out.append(type(e).__name__)
    finally:
        q.put((rank, out))
        dist.barrier()
        dist.destroy_process_group()
        be.close()


@pytest.mark.parametrize("world", [2, 3])
def test_streamed_exchange_host_logic_equals_oracle(world, tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    lock = str(tmp_path / "atomics.lock")
    procs = [ctx.Process(target=_stream_worker, args=(r, world, port, lock, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = []
    for ci, (mode, n, p_ood, p_ign) in enumerate(CASES[:4]):
        s, l = gi.metric_case(100 + ci, n, mode, p_ood, p_ign)
        w = c_oracle.eval_ood_measure(s, l)
        want += [w, w, ("refined", True)]
    want += [("copies", True), "MssError"]
    for r in range(world):
        assert results[r] == want, (r, results[r], want)
