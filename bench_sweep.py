#!/usr/bin/env python
"""bench_sweep.py -- BASELINE.json configs[3]: RoadAnomaly/SMIYC-shaped evaluation sweep.

    python bench_sweep.py [--images 2000] [--steps 3] [--warmup 1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_sweep.py --gpus N

One "step" = the whole tester loop of test_deeplab.py:84-117 for a dataset of `--images` synthetic
1024 x 2048 frames: per batch, fused DeepLab energy scoring + ignore masking + key append (one kernel, no
D2H), then ONE exact tie-aware AUROC / AP / FPR@95 over all valid pixels of the dataset.  Images are sharded
over the ranks by batch; the global metric uses the key-range exchange of evaluator.StreamingEvaluator
(all-reduce of counts and of a 2^16-bin key histogram, all-to-all of (key, label) by key range, all-gather
of the per-threshold integer counts), so the N-GPU result is bit-identical to the 1-GPU result.

Dataset (same for every world size): every batch is the same 16-image pool (logits as in bench.py cfg-1:
randn, x0.5 on OOD pixels, x2 elsewhere; labels 90/5/5 % ID/OOD/ignore), so the dataset is `images/16`
copies of the pool -- cross-image ties exactly as SURVEY 8(d) describes.  Replicating a dataset k times
multiplies every integer count by k and leaves every float64 ratio, hence AUROC/AP/FPR95, bit-identical,
so rank 0 also evaluates the pool ONCE and checks the sweep result against it (`matches_single_pool`).

Prints ONE JSON line (rank 0): value = dataset images / second (whole job, max over ranks), plus Mpix/s,
the per-phase split and the float64 results as hex so runs at N = 1, 2, 4, 8 can be compared bit for bit.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POOL, C, H, W = 16, 19, 1024, 2048


def make_pool(device):
    g = torch.Generator(device=device).manual_seed(4000)          # same pool on every rank
    x = torch.randn((POOL, C, H, W), device=device, generator=g)
    r = torch.rand((POOL, H, W), device=device, generator=g)
    lab = torch.zeros((POOL, H, W), dtype=torch.uint8, device=device)
    lab[r < 0.05] = 1
    lab[r > 0.95] = 255
    ood = (lab == 1).unsqueeze(1)
    x = torch.where(ood, 0.5 * x, 2.0 * x)
    return x.contiguous(), lab.contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--images", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench_sweep.py needs a B200: there is no CPU fallback for the product path")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif args.gpus > 1:
        raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")

    from multishiftseg_b200 import _lib as L
    from multishiftseg_b200.evaluator import StreamingEvaluator
    dev = torch.device("cuda", local)
    logits, labels = make_pool(dev)
    batches = max(1, args.images // POOL)
    images = batches * POOL
    my_batches = len(range(rank, batches, world))                   # batch j -> rank j % world
    ev = StreamingEvaluator(capacity=max(my_batches, 1) * POOL * H * W, device=dev, distributed=(world > 1))

    def sync_all():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        ev.reset()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        for _ in range(my_batches):
            ev.update_from_logits(logits, labels, key="energy", which=("energy",))
        e[1].record()
        res = ev.compute()
        e[2].record()
        torch.cuda.synchronize()
        return res, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])

    for _ in range(max(args.warmup, 1)):
        step()
    times, score_ms, metric_ms, res = [], [], [], None
    l0 = L.launch_count()
    for _ in range(args.steps):
        sync_all()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        res, a, b = step()
        t1.record()
        torch.cuda.synchronize()
        t = torch.tensor([t0.elapsed_time(t1), a, b], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t[0])); score_ms.append(float(t[1])); metric_ms.append(float(t[2]))
    launches = L.launch_count() - l0
    if dist is not None:
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms = sum(times) / len(times)

    if rank == 0:
        # the pool once, single GPU, same kernels: replication leaves the float64 results bit-identical
        single = StreamingEvaluator(capacity=POOL * H * W, device=dev, distributed=False)
        single.update_from_logits(logits, labels, key="energy", which=("energy",))
        ref = single.compute()
        hexes = [float(v).hex() for v in res]
        line = {
            "metric": "eval images/s (cfg4: fused DeepLab energy scoring + exact AUROC/AP/FPR95 over the whole dataset)",
            "value": images / ms * 1e3, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "dtype": "f32 scores, u32 keys, i64 counts, f64 tail", "data": "synthetic",
            "config": {"workload": f"cfg4: {images} images 1024x2048 (= {batches} x the 16-image pool), "
                                   "energy score, labels 90/5/5 % ID/OOD/ignore",
                       "parallelism": f"batches sharded over {world} GPU(s); one key-range exchange for the global metric"},
            "mpix_s": images * H * W / ms / 1e3,
            "phases_ms": {"score_and_append": sum(score_ms) / len(score_ms), "global_metric": sum(metric_ms) / len(metric_ms)},
            "result": [float(v) for v in res], "result_hex": hexes,
            "matches_single_pool": [float(v).hex() for v in ref] == hexes,
            "gpu_launches": launches,
        }
        if world > 1 and getattr(ev, "last_exchange", None):
            x = ev.last_exchange
            line["exchange"] = {"recv_pairs_rank0": int(sum(x["recv_counts"])), "thresholds_per_rank": x["thresholds_per_rank"],
                                "phase_ms_rank0": {k: round(v, 3) for k, v in x.get("phase_ms", {}).items()}}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
