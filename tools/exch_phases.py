"""Per-phase cycle counts of the remote-append exchange tile (development aid).  Needs the profiling build of the library:
    MSS_NVCC_EXTRA=-DMSS_EXCH_PROFILE python -m multishiftseg_b200.build --force
    cp multishiftseg_b200/libmss_b200.so multishiftseg_b200/libmss_b200_prof.so ; python -m multishiftseg_b200.build --force
    torchrun --nproc-per-node N tools/exch_phases.py [keys per rank]
"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multishiftseg_b200 import _lib as L  # noqa: E402

L.LIB_PATH = os.path.join(ROOT, "multishiftseg_b200", "libmss_b200_prof.so")
from multishiftseg_b200.evaluator import StreamingEvaluator  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
    g = torch.Generator(device="cuda").manual_seed(10 + rank)
    lib = L.load()
    lib.mss_debug_exchange_profile.restype = C.c_int
    lib.mss_debug_exchange_profile.argtypes = [C.c_void_p, C.c_int]
    names = {0: "load keys", 1: "digits + rank", 2: "barrier 1", 3: "reserve (+barrier)", 4: "offsets + scatter (3 barriers)",
             6: "store issue (tid 0)", 7: "whole tile"}
    for rep in range(2):
        ev = StreamingEvaluator(n + 1024, distributed=True, exchange="p2p")
        for _ in range(4):
            s = torch.randn(n // 4, device="cuda", generator=g)
            lab = (torch.rand(n // 4, device="cuda", generator=g) < 0.05).to(torch.uint8)
            ev.update(s, lab)
        del s, lab
        lib.mss_debug_exchange_profile(None, 1)
        torch.cuda.synchronize()
        r = ev.compute()
        out = (C.c_uint64 * 32)()
        lib.mss_debug_exchange_profile(out, 0)
        if rank == 0 and rep == 1:
            tiles = max(int(out[15]), 1)
            print(f"world {world}: {n} keys per rank, {tiles} tiles, exchange phase "
                  f"{ev.last_exchange.get('phase_ms', {}).get('exchange_append_p2p')} ms, result {tuple(float(v) for v in r)}")
            for i, nm in names.items():
                print(f"  {nm:34s} {out[i] / tiles:9.0f} cycles per tile")
            print("  reservation round trip by destination:", [round(out[16 + d] / tiles) for d in range(world)])
            print("  bulk read-completion by destination:  ", [round(out[24 + d] / tiles) for d in range(world)])
        del ev
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
