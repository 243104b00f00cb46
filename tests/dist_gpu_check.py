"""Launched by torchrun (one process per GPU, NCCL): multi-GPU exact metrics must be bit-identical to the
1-GPU result and to the oracle.  Exit code 0 on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import gen_inputs as gi  # noqa: E402
from multishiftseg_b200 import metric  # noqa: E402
from multishiftseg_b200.evaluator import StreamingEvaluator  # noqa: E402
from oracle import c_oracle  # noqa: E402


def make_case(seed, n, mode, p_ood, p_ign):
    if mode != "narrow":
        return gi.metric_case(seed, n, mode, p_ood, p_ign, label_dtype="uint8")
    # every score inside ONE top-16-bit key bin (0.999 .. 1.0): bin-granular splitters would send everything to one rank
    rng = np.random.default_rng(seed)
    s = (0.999 + 0.001 * rng.random(n)).astype(np.float32)
    r = rng.random(n)
    l = np.where(r < p_ood, 1, np.where(r > 1 - p_ign, 255, 0)).astype(np.uint8)
    return s, l


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    cases = [("cont", 2_000_003, 0.05, 0.05), ("q2", 1_000_000, 0.05, 0.2), ("f16", 3_000_001, 0.01, 0.05),
             ("const", 50_000, 0.3, 0.0), ("zeros", 90_000, 0.3, 0.05), ("cont", 1000, 0.2, 0.0),
             ("narrow", 2_000_000, 0.05, 0.05)]
    for ci, (mode, n, p_ood, p_ign) in enumerate(cases):
        s, l = make_case(500 + ci, n, mode, p_ood, p_ign)
        img = max(n // 16, 1)
        chunks = [(i, min(i + img, n)) for i in range(0, n, img)]
        want = c_oracle.eval_ood_measure(s, l)
        one = metric.eval_ood_measure(s, l)
        one = None if one is None else tuple(float(v) for v in one)
        for exch in ("p2p", "p2p_counted", "nccl"):   # remote append, counted peer-memory scatter, partition + NCCL all-to-all
            ev = StreamingEvaluator(n // world + 2 * img, distributed=True, exchange=exch)
            for a, b in chunks[rank::world]:
                ev.update(torch.from_numpy(s[a:b]).cuda(), torch.from_numpy(l[a:b]).cuda())
            got = ev.compute()
            got = None if got is None else tuple(float(v) for v in got)
            good = (got == want == one)
            if good and mode == "narrow" and world > 1:        # the second-level histogram must have balanced the key ranges
                rc = ev.last_exchange.get("dst_totals")
                good = bool(ev.last_exchange.get("refined_bins")) and max(rc) <= 1.25 * (sum(rc) / world) + 4096
            ok &= good
            if rank == 0:
                x = ev.last_exchange if got else {}
                print(f"[world {world}] {mode:6s} n={n:8d} {exch:11s}->{x.get('exchange')} multi-gpu == 1-gpu == oracle: {good}  {got}  "
                      f"dst_totals={x.get('dst_totals', '')} {x.get('p2p_error') or ''}", flush=True)
    # streaming exchange: every batch goes to its owners right behind its append (key ranges fixed from the first batch)
    for ci, (mode, n, p_ood, p_ign) in enumerate(cases[:4]):
        s, l = gi.metric_case(500 + ci, n, mode, p_ood, p_ign, label_dtype="uint8")
        img = max(n // 16, 1)
        chunks = [(i, min(i + img, n)) for i in range(0, n, img)]
        want = c_oracle.eval_ood_measure(s, l)
        for exch in ("stream", "stream_sm"):                       # copy-engine staged / remote-append kernel on a side stream
            ev = StreamingEvaluator(n // world + 2 * img, distributed=True, exchange=exch, stage_capacity=img + 16)
            for rep in range(2):                               # a second round after reset() reuses ranges and buffers
                ev.reset()
                for a, b in chunks[rank::world]:
                    ev.update(torch.from_numpy(s[a:b]).cuda(), torch.from_numpy(l[a:b]).cuda())
                got = ev.compute()
                got = None if got is None else tuple(float(v) for v in got)
                good = got == want
                ok &= good
                if rank == 0:
                    print(f"[world {world}] {mode:6s} n={n:8d} {exch} round {rep} multi-gpu == oracle: {good}  {got}", flush=True)
            del ev
    # fused DeepLab scoring -> evaluator, sharded images
    g = torch.Generator().manual_seed(7)
    B, H, W = 4, 128, 256
    x = torch.randn((B, 19, H, W), generator=g)
    r = torch.rand((B, H, W), generator=g)
    lab = torch.where(r < 0.05, 1, torch.where(r > 0.95, 255, 0)).to(torch.uint8)
    x = torch.where((lab == 1).unsqueeze(1), 0.5 * x, 2.0 * x)
    ev = StreamingEvaluator(B * H * W, distributed=True)
    for b in range(rank, B, world):
        ev.update_from_logits(x[b:b + 1].cuda(), lab[b:b + 1].cuda(), key="energy")
    got = tuple(float(v) for v in ev.compute())
    from multishiftseg_b200 import deeplab
    e = deeplab.energy_func(x.cuda())
    one = tuple(float(v) for v in metric.eval_ood_measure(e, lab.cuda()))
    want = c_oracle.eval_ood_measure(e.cpu().numpy(), lab.numpy())
    good = got == one == want
    ok &= good
    if rank == 0:
        print(f"[world {world}] fused deeplab->evaluator: {good} {got}", flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
