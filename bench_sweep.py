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


class Env:
    """Process / device context shared by bench.py and bench_sweep.py (one process per GPU)."""

    def __init__(self, dist, rank, world, local):
        self.dist, self.rank, self.world, self.local = dist, rank, world, local
        self.dev = torch.device("cuda", local)

    def sync_all(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = torch.tensor(vals, device=self.dev, dtype=torch.float64)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def sum_over_ranks(self, v):
        t = torch.tensor([v], device=self.dev, dtype=torch.int64)
        if self.dist is not None:
            self.dist.all_reduce(t)
        return int(t.item())


def _timed_steps(env, step, steps, warmup):
    """`step()` -> (result, score_ms, metric_ms).  Device-timed, max over ranks per step, mean over steps."""
    from multishiftseg_b200 import _lib as L
    for _ in range(max(warmup, 1)):
        step()
    times, score_ms, metric_ms, res = [], [], [], None
    l0 = L.launch_count()
    for _ in range(steps):
        env.sync_all()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        res, a, b = step()
        t1.record()
        torch.cuda.synchronize()
        t, a, b = env.max_over_ranks([t0.elapsed_time(t1), a, b])
        times.append(t); score_ms.append(a); metric_ms.append(b)
    launches = env.sum_over_ranks(L.launch_count() - l0)
    mean = lambda x: sum(x) / len(x)
    return res, mean(times), mean(score_ms), mean(metric_ms), launches


def run_cfg4(env, images=2000, steps=3, warmup=1, oracle_check=True, exchange="auto"):
    """cfg-4: `images` frames = images/16 copies of a 16-image pool, fused energy scoring + append per batch, ONE exact
    global metric.  Returns the result dict (rank 0; None elsewhere).  `bit_exact_vs_pool` is computed inside the run:
    sweep result == the pool evaluated once on one GPU == (oracle_check) the CPU oracle on the pool's score map."""
    from multishiftseg_b200.evaluator import StreamingEvaluator
    dev, rank, world = env.dev, env.rank, env.world
    logits, labels = make_pool(dev)
    batches = max(1, images // POOL)
    images = batches * POOL
    my_batches = len(range(rank, batches, world))                   # batch j -> rank j % world
    ev = StreamingEvaluator(capacity=max(my_batches, 1) * POOL * H * W, device=dev, distributed=(world > 1),
                            exchange=exchange if world > 1 else "auto")

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

    res, ms, score_ms, metric_ms, launches = _timed_steps(env, step, steps, warmup)
    exchange = getattr(ev, "last_exchange", None)
    del ev
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    # the pool once, single GPU, same kernels: replication leaves the float64 results bit-identical
    single = StreamingEvaluator(capacity=POOL * H * W, device=dev, distributed=False)
    energy = single.update_from_logits(logits, labels, key="energy", which=("energy",))["energy"]
    ref = single.compute()
    hexes = [float(v).hex() for v in res]
    pool_hex = [float(v).hex() for v in ref]
    out = {
        "workload": f"cfg4: {images} images 1024x2048 (= {batches} x the 16-image pool), fused energy scoring + append, "
                    "one exact AUROC/AP/FPR95 over all valid pixels; labels 90/5/5 % ID/OOD/ignore",
        "images": images, "n_gpus": world, "steps": steps, "scaling": "strong",
        "images_s": images / ms * 1e3, "mpix_s": images * H * W / ms / 1e3, "ms_per_step": ms,
        "phases_ms": {"score_and_append": score_ms, "global_metric": metric_ms},
        "result": [float(v) for v in res], "result_hex": hexes, "pool_result_hex": pool_hex,
        "matches_single_pool": pool_hex == hexes, "gpu_launches": launches,
    }
    if oracle_check:
        from oracle import c_oracle                                   # the checker, never the thing measured
        want = c_oracle.eval_ood_measure(energy.cpu().numpy(), labels.cpu().numpy())
        out["oracle_result_hex"] = [float(v).hex() for v in want]
        out["pool_matches_oracle"] = out["oracle_result_hex"] == pool_hex
        out["bit_exact_vs_pool"] = bool(out["matches_single_pool"] and out["pool_matches_oracle"])
    else:
        out["bit_exact_vs_pool"] = bool(out["matches_single_pool"])
    if world > 1 and exchange:
        out["exchange"] = {"kind": exchange["exchange"], "recv_keys_rank0": [int(sum(c)) for c in exchange["recv_counts"]],
                           "note": "phase_ms_rank0 are host wall-clock marks inside compute(); with kind = stream the first one "
                                   "also waits for the scoring / exchange kernels still in flight",
                           "thresholds_per_rank": exchange["thresholds_per_rank"],
                           "phase_ms_rank0": {k: round(v, 3) for k, v in exchange.get("phase_ms", {}).items()}}
    return out


def run_continuous(env, images=256, steps=3, warmup=1):
    """T ~= N variant of the sweep's metric stage: `images` frames of i.i.d. continuous scores (no replication, so
    nearly every valid pixel is its own threshold -- the worst case for the float64 tail), appended frame by frame and
    evaluated once.  Frame j is generated from seed 9000 + j whatever the world size, so result_hex is comparable
    across N."""
    from multishiftseg_b200.evaluator import StreamingEvaluator
    dev, rank, world = env.dev, env.rank, env.world
    mine = list(range(rank, images, world))
    ev = StreamingEvaluator(capacity=max(len(mine), 1) * H * W, device=dev, distributed=(world > 1))
    g = torch.Generator(device=dev)

    def frame(j):
        g.manual_seed(9000 + j)
        r = torch.rand(H * W, device=dev, generator=g)
        lab = torch.zeros(H * W, dtype=torch.uint8, device=dev)
        lab[r < 0.05] = 1
        lab[r > 0.95] = 255
        s = torch.randn(H * W, device=dev, generator=g) + (lab == 1) * 1.5
        return s, lab

    def fill():
        ev.reset()
        for j in mine:
            ev.update(*frame(j))

    def step():
        fill()                                                       # generation + append: not the measured part
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        res = ev.compute()
        e[1].record()
        torch.cuda.synchronize()
        return res, 0.0, e[0].elapsed_time(e[1])

    res, _, _, metric_ms, _ = _timed_steps(env, step, steps, warmup)
    exchange = getattr(ev, "last_exchange", None)
    fill()
    m = env.sum_over_ranks(ev.backend.state(ev.buf)[0])
    del ev
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    out = {"workload": f"{images} frames 1024x2048 of continuous i.i.d. scores (T ~= N): exact AUROC/AP/FPR95 of scores "
                       "already in the evaluator (metric stage only)",
           "valid_keys": m, "n_gpus": world, "metric_ms": metric_ms, "gkeys_s": m / metric_ms / 1e6,
           "gpix_s": images * H * W / metric_ms / 1e6, "result_hex": [float(v).hex() for v in res]}
    if world > 1 and exchange:
        out["thresholds"] = int(sum(exchange["thresholds_per_rank"]))
        out["phase_ms_rank0"] = {k: round(v, 3) for k, v in exchange.get("phase_ms", {}).items()}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--cfg", type=int, default=4, choices=(4, 5))
    ap.add_argument("--images", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--continuous", type=int, default=0, help="also run the T ~= N metric-stage variant on this many frames")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=("auto", "stream", "stream_sm", "p2p", "p2p_counted", "nccl"),
                    help="multi-GPU exchange of StreamingEvaluator (stream = overlapped with the scoring)")
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

    env = Env(dist, rank, world, local)
    r = run_cfg4(env, args.images, args.steps, args.warmup, oracle_check=not args.no_oracle, exchange=args.exchange)
    c = run_continuous(env, args.continuous, args.steps, args.warmup) if args.continuous else None
    if rank == 0:
        line = {
            "metric": "eval images/s (cfg4: fused DeepLab energy scoring + exact AUROC/AP/FPR95 over the whole dataset)",
            "value": r["images_s"], "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "dtype": "f32 scores, u32 keys, i64 counts, f64 tail", "data": "synthetic",
            "config": {"workload": r["workload"],
                       "parallelism": f"batches sharded over {world} GPU(s); one key-range exchange for the global metric"},
        }
        line.update({k: v for k, v in r.items() if k not in ("workload", "n_gpus", "steps", "scaling", "ms_per_step")})
        if c:
            line["continuous"] = c
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
