#!/usr/bin/env python
"""bench_sweep.py -- BASELINE.json configs[3] (RoadAnomaly/SMIYC-shaped evaluation sweep) and, with --cfg 5,
configs[4] (mixed multi-shift evaluation, 1080 x 1920, Mask2Former + DeepLab scoring, global metric merge).

    python bench_sweep.py [--cfg 4|5] [--images 2000] [--steps 3] [--warmup 1]
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

--cfg 5: two datasets (ACDC-POC- and MUAD-shaped, 1080 x 1920 frames, `--images` frames each, default
256), every frame scored by BOTH models' post-head paths --
  * DeepLab: OOD-head logits at half resolution [B, 19, 540, 960] -> energy -> align_corners=True upsample
    (deepv3.py:282-283) -> evaluator append;
  * Mask2Former: class logits [B, 100, 20] + decoder masks [B, 100, 272, 480] (frame padded to 1088 x 1920,
    size divisibility 32) -> fused upsample / sigmoid / contraction / 1 - max, cropped to 1080 x 1920
    (maskformer_model.py:271-277, train_m2f.py:387-407) -> evaluator append --
and ONE exact global metric per dataset over both models' (score, label) streams (the key-range exchange merges
the ranks' shards).  value = dataset frames / second, every frame scored by both models.
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


H5, W5, POOL5 = 1080, 1920, 8


def make_pool5(device, seed):
    """cfg-5 pool of one dataset: half-resolution DeepLab OOD-head logits, Mask2Former OOD-head outputs, labels with
    rectangular OOD blobs (1) and a 5 % ignore sprinkle (255)."""
    g = torch.Generator(device=device).manual_seed(seed)
    lab = torch.zeros((POOL5, H5, W5), dtype=torch.uint8, device=device)
    boxes = torch.randint(0, 10 ** 6, (POOL5, 3, 4), device=device, generator=g).tolist()
    for b in range(POOL5):
        for (a, c, d, e) in boxes[b]:
            y0, x0 = a % (H5 - 200), c % (W5 - 300)
            lab[b, y0:y0 + 40 + d % 160, x0:x0 + 60 + e % 240] = 1
    lab[torch.rand((POOL5, H5, W5), device=device, generator=g) > 0.95] = 255
    ood_half = (lab[:, ::2, ::2] == 1).unsqueeze(1)
    x = torch.randn((POOL5, C, H5 // 2, W5 // 2), device=device, generator=g)
    x = torch.where(ood_half, 0.5 * x, 2.0 * x).contiguous()
    cls = (3.0 * torch.randn((POOL5, 100, 20), device=device, generator=g)).contiguous()
    lo = torch.randn((POOL5, 100, 272 // 4, 480 // 4), device=device, generator=g)
    lo = (4.0 * torch.nn.functional.interpolate(lo, size=(272, 480), mode="bilinear", align_corners=False)).contiguous()
    return x, cls, lo, lab.contiguous()


def main5(args, world, rank, local, dist):
    from multishiftseg_b200 import _lib as L, deeplab, m2f
    from multishiftseg_b200.evaluator import StreamingEvaluator
    dev = torch.device("cuda", local)
    names = ("ACDC-POC-shaped", "MUAD-shaped")
    pools = [make_pool5(dev, 5000 + i) for i in range(len(names))]
    n_img = args.images if args.images != 2000 else 256
    batches = max(1, n_img // POOL5)
    images = batches * POOL5
    my_batches = len(range(rank, batches, world))
    ev = StreamingEvaluator(capacity=max(my_batches, 1) * POOL5 * H5 * W5 * 2, device=dev, distributed=(world > 1))

    def feed(e, pool):
        x, cls, lo, lab = pool
        e.update(deeplab.anomaly_score(x, (H5, W5)), lab)
        e.update(m2f.anomaly_score_from_lowres(cls, lo, (1088, W5), (H5, W5)), lab)

    def sync_all():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        out, score_ms, metric_ms = [], 0.0, 0.0
        for pool in pools:
            ev.reset()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            for _ in range(my_batches):
                feed(ev, pool)
            e[1].record()
            out.append(ev.compute())
            e[2].record()
            torch.cuda.synchronize()
            score_ms += e[0].elapsed_time(e[1]); metric_ms += e[1].elapsed_time(e[2])
        return out, score_ms, metric_ms

    for _ in range(max(args.warmup, 1)):
        step()
    times, score_ms, metric_ms, res = [], [], [], None
    l0 = L.launch_count()
    for _ in range(args.steps):
        sync_all()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
        single = StreamingEvaluator(capacity=POOL5 * H5 * W5 * 2, device=dev, distributed=False)
        refs = []
        for pool in pools:
            single.reset()
            feed(single, pool)
            refs.append(single.compute())
        hexes = [[float(v).hex() for v in r] for r in res]
        frames = images * len(names)
        line = {
            "metric": "eval images/s (cfg5: Mask2Former + DeepLab post-head scoring of every frame + exact AUROC/AP/FPR95 "
                      "per dataset over both models' streams)",
            "value": frames / ms * 1e3, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "dtype": "f32 scores, u32 keys, i64 counts, f64 tail", "data": "synthetic",
            "config": {"workload": f"cfg5: {len(names)} datasets x {images} frames 1080x1920 (= {batches} x an 8-frame pool), "
                                   "DeepLab half-res energy + align_corners upsample, M2F 272x480 masks padded to 1088x1920 "
                                   "and cropped, labels: OOD boxes + 5 % ignore",
                       "parallelism": f"batches sharded over {world} GPU(s); one key-range exchange per dataset metric"},
            "mpix_scored_s": 2 * frames * H5 * W5 / ms / 1e3,
            "phases_ms": {"score_and_append": sum(score_ms) / len(score_ms), "global_metric": sum(metric_ms) / len(metric_ms)},
            "result": {n: [float(v) for v in r] for n, r in zip(names, res)},
            "result_hex": {n: h for n, h in zip(names, hexes)},
            "matches_single_pool": [[float(v).hex() for v in r] for r in refs] == hexes,
            "gpu_launches": launches,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--cfg", type=int, default=4, choices=(4, 5))
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
    if args.cfg == 5:
        return main5(args, world, rank, local, dist)

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
