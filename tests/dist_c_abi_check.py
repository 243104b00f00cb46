"""Launched by torchrun (one process per GPU): the C entry point mss_ood_metrics_dist, called the way a C / C++ binder of
include/mss_b200.h would -- with its own ncclComm_t (created here through ctypes on the NCCL library of the process, the
unique id broadcast with torch.distributed) -- must return the oracle's result on every rank.  Exit code 0 on success."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import gen_inputs as gi  # noqa: E402
from multishiftseg_b200 import _lib as L, metric  # noqa: E402
from oracle import c_oracle  # noqa: E402


class UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


def make_comm(rank, world):
    nccl = C.CDLL("libnccl.so.2")
    nccl.ncclGetUniqueId.argtypes = [C.POINTER(UniqueId)]
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    uid = UniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    t = torch.frombuffer(bytearray(bytes(uid.internal)), dtype=torch.uint8).cuda()
    dist.broadcast(t, 0)
    C.memmove(uid.internal, bytes(t.cpu().numpy().tobytes()), 128)
    comm = C.c_void_p()
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    return nccl, comm


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    nccl, comm = make_comm(rank, world)
    lib = L.load()
    ok = True
    cases = [("cont", 2_000_003, 0.05, 0.05), ("q2", 1_000_000, 0.05, 0.2), ("f16", 3_000_001, 0.01, 0.05),
             ("const", 50_000, 0.3, 0.0), ("cont", 1000, 0.2, 0.0), ("onlyid", 5000, 0.0, 0.1),
             ("narrow", 2_000_000, 0.05, 0.05)]
    for ci, (mode, n, p_ood, p_ign) in enumerate(cases):
        if mode == "narrow":      # all scores in one top-16-bit key bin: the second-level histogram must split it
            rng = np.random.default_rng(600 + ci)
            s = (0.999 + 0.001 * rng.random(n)).astype(np.float32)
            r = rng.random(n)
            l = np.where(r < p_ood, 1, np.where(r > 1 - p_ign, 255, 0)).astype(np.uint8)
        else:
            s, l = gi.metric_case(600 + ci, n, "cont" if mode == "onlyid" else mode, max(p_ood, 0.01), p_ign, label_dtype="uint8")
        if mode == "onlyid":
            l[l == 1] = 0
        want = c_oracle.eval_ood_measure(s, l)
        img = max(n // 16, 1)
        chunks = [(i, min(i + img, n)) for i in range(0, n, img)]
        buf = metric.PairBuffer(n // world + 2 * img, "cuda")
        buf.reset()
        for a, b in chunks[rank::world]:
            buf.append(torch.from_numpy(s[a:b]).cuda(), torch.from_numpy(l[a:b]).cuda())
        out, cnt = (C.c_double * 3)(), (C.c_int64 * 4)()
        rc = lib.mss_ood_metrics_dist(C.byref(buf.c), comm, rank, world, out, cnt, torch.cuda.current_stream().cuda_stream)
        got = tuple(out) if rc == 0 else None
        good = (rc in (0, L.MSS_EMPTY_CLASS)) and got == want
        if good and want is not None:
            v = (l != 255)
            good = list(cnt)[:2] == [int((l == 1).sum()), int((l == 0).sum())] and cnt[2] == np.unique(s[v].astype(np.float32) + 0.0).size
        ok &= bool(good)
        if rank == 0:
            print(f"[world {world}] C ABI mss_ood_metrics_dist {mode:6s} n={n:8d} rc={rc} == oracle: {good}  {got} counts={list(cnt)}", flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
