#!/usr/bin/env python
"""bench.py -- headline benchmark of the scoring + OOD-evaluation path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input:
BASELINE.json configs[1] -- DeepLabv3+ scoring of a 16 x 19 x 1024 x 2048 fp32 logit batch into
max-logit + energy + entropy maps (88 B/px algorithmic: 76 read + 12 written).  Weak scaling: every
rank scores its own 16-image batch.  Prints ONE JSON line (rank 0).

  value      Mpix/s with the logits resident in HBM (device-timed with CUDA events, max over ranks)
  e2e        the same metric through the host-buffer C-ABI entry point (pinned host logits in, pinned
             host score maps out; H2D + D2H inside the timed region)
  roofline   achieved HBM GB/s of the scoring kernel vs MEASURED_PEAKS.json
  cpu_baseline  the oracle port (torch-CPU one-liners of deepv3.py:251-253 + extras) on the host cores
  extra.eval_sweep   (every N) BASELINE configs[3]: the 2000-image evaluation sweep -- fused scoring + append, ONE exact
             global metric through the key-range exchange -- with eval images/s, per-phase ms, the float64 results
             as hex and `bit_exact_vs_pool` (sweep == the pool on one GPU == the CPU oracle), so the driver's
             N = 1, 2, 4, 8 lines carry the real multi-GPU path and its strong scaling
  extra      (N = 1) the other kernels of the path (exact metrics with a CUB yardstick, Mask2Former fused inference,
             head GEMMs, ...), each with its own throughput / roofline fraction, and the CPU legs of BASELINE.md
             section 3 (cfg-1, cfg-3, cfg-4 subset) with GPU == CPU metric equality checked in the run

`--impl reference` times the reference's own CPU implementation of the path (oracle port: the reference is
plain PyTorch / numpy / scikit-learn, which is exactly what the port calls) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, C, H, W = 16, 19, 1024, 2048
WHICH = ("maxlogit", "energy", "entropy")
BYTES_PER_PX = 4 * C + 4 * len(WHICH)           # 88: SURVEY 8(d)
METRIC = "Mpix/s scored (DeepLabv3+ max-logit+energy+entropy, 16x19x1024x2048 fp32 per GPU)"
WORKLOAD = "cfg2: DeepLabv3+ scoring batch 16x19x1024x2048 fp32 -> max-logit + energy + entropy"


def common_config(world):
    """identical for the b200 and the reference arm (the driver compares them)"""
    return {"workload": WORKLOAD, "batch_per_gpu": B_PER_GPU, "classes": C, "frame": [H, W],
            "l2": "input 2.55 GB per step >> 126 MB L2, no flush needed",
            "parallelism": f"images sharded over {world} GPU(s), no data-path collective"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)", d
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", {}


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while `active` is set (the timed regions)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.power = [], set(), []
        self.active = threading.Event()
        self.stop = threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self.stop.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
            time.sleep(0.005)

    def summary(self):
        self.stop.set()
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.err}
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return dist, rank, world, local
    if n_gpus > 1:
        raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(0)
    return None, 0, 1, 0


def timed(fn, steps, dist, sampler=None):
    """K steps between barrier+sync, CUDA events on the launching (current) stream, max over ranks -> ms total."""
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler is not None:
        sampler.active.set()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.active.clear()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


# ----------------------------------------------------------------------------------------------------------
def cpu_reference_scoring(sample_images, min_seconds=10.0, max_seconds=40.0):
    """Oracle port of the scoring step on the host cores (torch CPU, all threads)."""
    from oracle import scoring_oracle as so
    g = torch.Generator().manual_seed(0)
    x = torch.randn((sample_images, C, H, W), generator=g)

    def step():
        return so.maxlogit_score(x), so.energy_func(x), so.entropy_score(x)

    step()
    t0, reps = time.perf_counter(), 0
    while True:
        step()
        reps += 1
        el = time.perf_counter() - t0
        if el >= min_seconds or el >= max_seconds:
            break
    mpix = reps * sample_images * H * W / el / 1e6
    return mpix, reps, el


def run_reference(args):
    """`--impl reference`: the reference's own CPU path (oracle port) for the same metric / config: every step scores
    the whole 16-image batch (in 2-image slices, to bound the fp32 temporaries of the eager expressions)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import scoring_oracle as so
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn((B_PER_GPU, C, H, W), generator=g)

    def step():
        out = []
        for b in range(0, B_PER_GPU, 2):
            xs = x[b:b + 2]
            out.append((so.maxlogit_score(xs), so.energy_func(xs), so.entropy_score(xs)))
        return out

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    mpix = args.steps * B_PER_GPU * H * W / el / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(args.gpus),
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} steps x the whole 16-image batch (1024x2048), torch-CPU restatement of "
                                   "deepv3.py:251-253 + max-logit + entropy (oracle/scoring_oracle.py); host threads only, "
                                   "one batch whatever --gpus says"},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
def extra_metrics_stage(images=4):
    """Exact AUROC/AP/FPR95 on `images` x 1024 x 2048 score/label maps resident in HBM, the sort alone at three sizes
    beside cub::DeviceRadixSort on the same box, and the one-image latency."""
    from multishiftseg_b200 import metric
    peak, _, _ = peaks()
    g = torch.Generator(device="cuda").manual_seed(4000)

    def case(n):
        lab = torch.zeros(n, dtype=torch.uint8, device="cuda")
        r = torch.rand(n, device="cuda", generator=g)
        lab[r < 0.05] = 1
        lab[r > 0.95] = 255
        s = torch.randn(n, device="cuda", generator=g) + (lab == 1) * 1.5
        return s, lab

    def sort_ms(s, lab):
        buf = metric.PairBuffer(s.numel(), "cuda")
        buf.append(s, lab)
        m = buf.read_state()[0]
        keys0 = buf.keys.clone()
        ts = []
        for _ in range(6):
            buf.keys.copy_(keys0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            buf.sort(m)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts[1:]), m

    n = images * H * W
    s, lab = case(n)
    for _ in range(2):
        res = metric.eval_ood_measure(s, lab)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        res = metric.eval_ood_measure(s, lab)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    sm, m = sort_ms(s, lab)
    out = {"workload": f"exact AUROC/AP/FPR95, {images}x1024x2048 px in HBM (90/5/5 % ID/OOD/ignore, continuous scores: T ~= N)",
           "mpix_s": n / t / 1e6, "images_s": images / t, "ms": t * 1e3, "valid_keys": m,
           "result": [float(x) for x in res],
           "sort": {"ms": sm, "gkeys_s": m / sm / 1e6,
                    "algorithmic_GBs": m * 4 / sm / 1e6,       # 4 B/key read once (the label is the stream, not a byte)
                    "implementation_GBs": m * 36 / sm / 1e6,   # 4 + 4 x (4 + 4) B/key actually moved
                    "frac_of_hbm_peak_impl": m * 36 / sm / 1e6 / peak}}
    del s, lab
    # one image (what test_deeplab.py evaluates per small dataset) and the larger sizes
    sizes = [H * W, 16 * H * W, 64 * H * W]
    rows = []
    for nn in sizes:
        s, lab = case(nn)
        metric.eval_ood_measure(s, lab)
        torch.cuda.synchronize()
        tt = []
        for _ in range(5):
            t0 = time.perf_counter()
            metric.eval_ood_measure(s, lab)
            torch.cuda.synchronize()
            tt.append(time.perf_counter() - t0)
        sm2, m2 = sort_ms(s, lab)
        rows.append({"px": nn, "eval_ood_measure_ms": statistics.median(tt) * 1e3, "gpix_s": nn / statistics.median(tt) / 1e9,
                     "valid_keys": m2, "sort_ms": sm2, "sort_gkeys_s": m2 / sm2 / 1e6})
        del s, lab
    out["by_size"] = rows
    torch.cuda.empty_cache()
    out["sort"]["cub_yardstick"] = cub_yardstick([2 * H * W, 4 * H * W, 16 * H * W, 64 * H * W])
    return out


def extra_m2f(batch=8):
    """cfg-3: Mask2Former fused post-head inference, Q=100, C=19+1, 256x512 -> 1024x2048, batch 8."""
    from multishiftseg_b200 import m2f
    g = torch.Generator(device="cuda").manual_seed(3000)
    cls = 3.0 * torch.randn((batch, 100, 20), device="cuda", generator=g)
    lo = 4.0 * torch.randn((batch, 100, 256, 512), device="cuda", generator=g)
    out = {}
    for name, fn in (("anomaly_score", lambda: m2f.anomaly_score_from_lowres(cls, lo, (H, W), (H, W))),
                     ("semseg19", lambda: m2f.post_head_inference(cls, lo, (H, W), extra_channels=False))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        px = batch * H * W
        flop = px * (2 * 100 * 19)                       # contraction only (3 800 flop/px)
        out[name] = {"ms": ms, "mpix_s": px / ms / 1e3, "images_s": batch / ms * 1e3,
                     "contraction_TFLOPs": flop / ms / 1e9}
    # the reference evaluates one image per call (exps/M2F.yaml:21, valid_batch 1): 1056 CTAs on 296 slots = 3.6 waves
    cls1, lo1 = cls[:1].contiguous(), lo[:1].contiguous()
    for _ in range(3):
        m2f.anomaly_score_from_lowres(cls1, lo1, (H, W), (H, W))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m2f.anomaly_score_from_lowres(cls1, lo1, (H, W), (H, W))
    e1.record()
    torch.cuda.synchronize()
    out["anomaly_score_batch1"] = {"us_per_image": e0.elapsed_time(e1) / 20 * 1e3, "launches_per_call": 2,
                                   "note": "device time of back-to-back calls (class-probability table kernel + main kernel)"}
    out["workload"] = f"cfg3: M2F semantic_inference Q=100 C=19+1 256x512->1024x2048 batch {batch}"
    out["fma_roofline_TFLOPs"] = 148 * 128 * 2 * 1.965e9 / 1e12
    # binding pipe of the tcgen05 kernel: the two MUFU ops (ex2 + rcp) of each of the Q sigmoids per pixel,
    # 16 lanes/clk/SM -> 2 * Q * px / (16 * 148 * f_SM)
    floor_ms = 2 * 100 * batch * H * W / (16 * 148 * 1.965e9) * 1e3
    out["mufu_floor_ms"] = floor_ms
    out["frac_of_mufu_roofline"] = floor_ms / out["anomaly_score"]["ms"]
    return out


def extra_m2f_chain():
    """SURVEY 8f-1, Mask2Former half: is fusing the mask-logit GEMM INTO the semantic-inference kernel worth it?  The
    two-launch chain writes the decoder-resolution masks once (52 MB per image) and reads them back.  Measured here: the
    chain per image at batch 8 (intermediate through HBM: 420 MB per batch > L2) and at batch 1 (the 52 MB intermediate
    stays in the 126 MB L2 -- the traffic a fused kernel would save is already off the DRAM path), beside its two parts."""
    from multishiftseg_b200 import m2f
    g = torch.Generator(device="cuda").manual_seed(7100)
    out = {}

    def ms_of(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for B in (8, 1):
        feat = torch.randn((B, 256, 256, 512), device="cuda", generator=g)
        emb = torch.randn((B, 100, 256), device="cuda", generator=g) / 16
        cls = 3.0 * torch.randn((B, 100, 20), device="cuda", generator=g)
        lo = m2f.mask_logits(emb, feat)
        t_gemm = ms_of(lambda: m2f.mask_logits(emb, feat))
        t_sem = ms_of(lambda: m2f.anomaly_score_from_lowres(cls, lo, (H, W), (H, W)))
        t_chain = ms_of(lambda: m2f.anomaly_score_from_features(cls, emb, feat, (H, W), (H, W)))
        out[f"batch{B}"] = {"mask_gemm_us_per_image": t_gemm * 1e3 / B, "semantic_us_per_image": t_sem * 1e3 / B,
                            "chain_us_per_image": t_chain * 1e3 / B,
                            "intermediate_MB": B * 100 * 256 * 512 * 4 / 1e6}
        del feat, emb, cls, lo
    b8, b1 = out["batch8"], out["batch1"]
    # what removing the intermediate's DRAM round trip can be worth at most: its bytes (written + read) at the HBM peak
    peak, _, _ = peaks()
    out["intermediate_round_trip_us_per_image_at_hbm_peak"] = 2 * 100 * 256 * 512 * 4 / peak / 1e3
    out["chain_minus_parts_us_per_image_batch8"] = b8["chain_us_per_image"] - b8["mask_gemm_us_per_image"] - b8["semantic_us_per_image"]
    out["workload"] = "anomaly_score_from_features: mask_embed x mask_features -> masks 256x512 -> fused upsample/sigmoid/contraction/1-max 1024x2048"
    return out


def extra_confusion():
    """mIoU confusion histogram fused with argmax over the cfg-2 logit batch (76 + 1 B/px read)."""
    from multishiftseg_b200 import segmetric
    g = torch.Generator(device="cuda").manual_seed(5000)
    x = torch.randn((B_PER_GPU, C, H, W), device="cuda", generator=g)
    gt = torch.randint(0, C, (B_PER_GPU, H, W), device="cuda", generator=g, dtype=torch.int64).to(torch.uint8)
    acc = segmetric.ConfusionAccumulator(C, "cuda")

    def timed_ms(xx, gg):
        for _ in range(3):
            acc.update_from_logits(xx, gg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            acc.update_from_logits(xx, gg)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10

    ms = timed_ms(x, gt)
    px = B_PER_GPU * H * W
    peak, _, _ = peaks()
    gbs = px * (4 * C + 1) / ms / 1e6
    out = {"workload": "fused argmax + 19x19 confusion histogram, 16x19x1024x2048 fp32 logits + u8 gt "
                       "(i.i.d. random gt and logits: every lane of a warp hits a different bin, the worst case)",
           "ms": ms, "mpix_s": px / ms / 1e3, "algorithmic_GBs": gbs, "frac_of_hbm_peak": gbs / peak}
    # segmentation-shaped input: 64x64-pixel regions of one class, the logits favour the region's class
    region = torch.randint(0, C, (B_PER_GPU, H // 64, W // 64), device="cuda", generator=g, dtype=torch.int64)
    gt2 = region.repeat_interleave(64, dim=1).repeat_interleave(64, dim=2).to(torch.uint8).contiguous()
    del region
    for c in range(C):                                            # in place: no second 2.5 GB logit tensor
        x[:, c].add_((gt2 == c).to(torch.float32), alpha=6.0)
    ms2 = timed_ms(x, gt2)
    gbs2 = px * (4 * C + 1) / ms2 / 1e6
    out["coherent"] = {"workload": "same shapes, 64x64-pixel regions of one class (gt) with logits favouring it",
                       "ms": ms2, "mpix_s": px / ms2 / 1e3, "algorithmic_GBs": gbs2, "frac_of_hbm_peak": gbs2 / peak}
    return out


def extra_head():
    """fused DeepLab head (two 1x1 convs + energy), 8 x 256 x 512 x 1024 fp32 features."""
    from multishiftseg_b200 import deeplab
    B, K, h, w = 8, 256, 512, 1024
    g = torch.Generator(device="cuda").manual_seed(6000)
    feat = torch.relu(torch.randn((B, K, h, w), device="cuda", generator=g))
    wc = torch.randn((C, K), device="cuda", generator=g) / 16
    wo = torch.randn((C, K), device="cuda", generator=g) / 16
    for _ in range(3):
        deeplab.head_scores(feat, wc, wo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        deeplab.head_scores(feat, wc, wo)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    px = B * h * w
    peak, _, _ = peaks()
    gbs = px * (K * 4 + C * 4 + 4) / ms / 1e6
    return {"workload": "fused ood_head/final[-1] 1x1 convs + energy, 8x256x512x1024 fp32 features (tcgen05 3xTF32)",
            "ms": ms, "head_mpix_s": px / ms / 1e3, "algorithmic_GBs": gbs, "frac_of_hbm_peak": gbs / peak}


def extra_mask_gemm():
    """Mask2Former mask-logit GEMM einsum("bqc,bchw->bqhw"), 8 x [100 x 256] x [256 x 256 x 512]."""
    from multishiftseg_b200 import m2f
    B, Q, K, h, w = 8, 100, 256, 256, 512
    g = torch.Generator(device="cuda").manual_seed(7000)
    feat = torch.randn((B, K, h, w), device="cuda", generator=g)
    emb = torch.randn((B, Q, K), device="cuda", generator=g) / 16
    for _ in range(3):
        m2f.mask_logits(emb, feat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m2f.mask_logits(emb, feat)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    px = B * h * w
    peak, _, _ = peaks()
    gbs = px * (K * 4 + Q * 4) / ms / 1e6
    return {"workload": "mask_embed x mask_features -> decoder-resolution mask logits, 8x[100x256]x[256x256x512] fp32 "
                        "(tcgen05 3xTF32)",
            "ms": ms, "us_per_image": ms * 1e3 / B, "algorithmic_GBs": gbs, "frac_of_hbm_peak": gbs / peak,
            "fp32_equivalent_TFLOPs": 2.0 * px * Q * K / ms / 1e9}


def extra_backward():
    """backward of energy_func (train_deeplab.py:197-198 differentiates through the scoring path)."""
    from multishiftseg_b200 import deeplab
    g = torch.Generator(device="cuda").manual_seed(8000)
    x = torch.randn((B_PER_GPU, C, H, W), device="cuda", generator=g).requires_grad_(True)
    go = torch.randn((B_PER_GPU, H, W), device="cuda", generator=g)
    s = deeplab.energy_func(x)
    for _ in range(3):
        torch.autograd.grad(s, x, go, retain_graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.autograd.grad(s, x, go, retain_graph=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    px = B_PER_GPU * H * W
    peak, _, _ = peaks()
    gbs = px * (2 * C * 4 + 4) / ms / 1e6
    return {"workload": "energy_func backward (-softmax * grad), 16x19x1024x2048 fp32 logits", "ms": ms,
            "mpix_s": px / ms / 1e3, "algorithmic_GBs": gbs, "frac_of_hbm_peak": gbs / peak}


def _best_of(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


def _gpu_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts), r


def cpu_legs():
    """the reference's CPU path (oracle port: the same torch / numpy / scikit-learn calls) timed on
    this box's host cores beside the GPU path, and the metric triples compared for EQUALITY in the run."""
    import numpy as np
    import bench_sweep
    from multishiftseg_b200 import deeplab, m2f, metric
    from oracle import metrics_oracle as mo, scoring_oracle as so
    hexes = lambda r: None if r is None else [float(v).hex() for v in r]
    out = {"cores": os.cpu_count(), "torch_threads": torch.get_num_threads(),
           "note": "CPU = oracle port of the reference flow (torch-CPU scoring with all threads; numpy + scikit-learn "
                   "metrics, single-threaded as in the reference); wall clock, best of 3 where repeated"}
    px = H * W
    # ---- cfg-1: one image, energy + AUROC/AP/FPR95 (SURVEY 8d synthetic input) ---------------------------------
    g = torch.Generator().manual_seed(0)
    x = torch.randn((1, C, H, W), generator=g)
    r = torch.rand((1, H, W), generator=g)
    lab = torch.where(r < 0.05, 1, torch.where(r > 0.95, 255, 0)).to(torch.int64)
    x = torch.where((lab == 1).unsqueeze(1), 0.5 * x, 2.0 * x)
    t_score, e_cpu = _best_of(lambda: so.energy_func(x))
    conf, labn = e_cpu.numpy(), lab.numpy()                        # float32 / int64, as test_deeplab.py:98-101 passes them
    t_metric, ref = _best_of(lambda: mo.eval_ood_measure(conf, labn), reps=1)
    xg, lg = x.cuda(), lab.cuda()
    g_score_ms, e_gpu = _gpu_ms(lambda: deeplab.energy_func(xg))
    g_metric_ms, got_dev = _gpu_ms(lambda: metric.eval_ood_measure(e_gpu, lg))
    # the call the reference makes: numpy float32 scores + numpy int64 labels on the HOST, result back on the host
    g_e2e_ms, got_host = _gpu_ms(lambda: metric.eval_ood_measure(conf, labn))
    out["cfg1"] = {
        "workload": "cfg1: 1x19x1024x2048 logits -> energy -> exact AUROC/AP/FPR95 (90/5/5 % ID/OOD/ignore)",
        "cpu": {"score_s": t_score, "score_mpix_s": px / t_score / 1e6, "metric_s": t_metric,
                "metric_mpix_s": px / t_metric / 1e6, "result_hex": hexes(ref)},
        "gpu": {"score_ms": g_score_ms, "score_mpix_s": px / g_score_ms / 1e3, "metric_ms_device_inputs": g_metric_ms,
                "metric_mpix_s_device_inputs": px / g_metric_ms / 1e3,
                "metric_e2e_ms_host_numpy_inputs": g_e2e_ms, "metric_e2e_mpix_s": px / g_e2e_ms / 1e3,
                "metric_e2e_h2d_bytes": conf.nbytes + labn.nbytes, "metric_e2e_d2h_bytes": 24,
                "result_hex_same_score_map": hexes(got_host), "result_hex_gpu_score_map": hexes(got_dev)},
        "metric_equal_given_equal_score_maps": hexes(got_host) == hexes(ref),
        "score_max_rel_err": float(((e_gpu.cpu() - e_cpu).abs() / e_cpu.abs().clamp_min(1e-6)).max()),
        "metric_stage_speedup_e2e": t_metric * 1e3 / g_e2e_ms,
    }
    del xg, lg, e_gpu
    # ---- cfg-3: Mask2Former post-head inference, one image ------------------------------------------------------
    g = torch.Generator().manual_seed(3000)
    cls = 3.0 * torch.randn((1, 100, 20), generator=g)
    lo = 4.0 * torch.randn((1, 100, 256, 512), generator=g)
    t_m2f, a_cpu = _best_of(lambda: so.m2f_anomaly_from_lowres(cls, lo, (H, W), (H, W)), reps=2)
    clsg, log_ = cls.cuda(), lo.cuda()
    g_m2f_ms, a_gpu = _gpu_ms(lambda: m2f.anomaly_score_from_lowres(clsg, log_, (H, W), (H, W)))
    out["cfg3"] = {
        "workload": "cfg3 (one image of the batch): Q=100, C=19+1, masks 256x512 -> 1024x2048, 1 - max_c score",
        "cpu": {"s_per_image": t_m2f, "mpix_s": px / t_m2f / 1e6},
        "gpu": {"ms_per_image_batch1": g_m2f_ms, "mpix_s": px / g_m2f_ms / 1e3},
        "allclose_rtol1e-5_atol2e-6": bool(torch.allclose(a_gpu.cpu(), a_cpu, rtol=1e-5, atol=2e-6)),
    }
    del clsg, log_, a_gpu
    # ---- cfg-4 / cfg-5 metric stage on a SUBSET: 8 images of the sweep's pool (the full 2000 images would take hours
    #      and > 150 GB of host RAM on the reference path) -----------------------------------------------------------
    sub = 8
    logits, labels = bench_sweep.make_pool(torch.device("cuda"))
    e8 = deeplab.energy_func(logits[:sub])
    l8 = labels[:sub]
    del logits
    conf8, lab8 = e8.cpu().numpy().reshape(-1), l8.cpu().numpy().astype(np.int64).reshape(-1)
    t_m8, ref8 = _best_of(lambda: mo.eval_ood_measure(conf8, lab8), reps=1)
    g_m8_ms, got8 = _gpu_ms(lambda: metric.eval_ood_measure(e8, l8))
    g_m8_e2e_ms, got8h = _gpu_ms(lambda: metric.eval_ood_measure(conf8, lab8), reps=3)
    out["cfg4_subset"] = {
        "workload": f"cfg4/cfg5 metric stage on a subset: {sub} of the sweep pool's images ({sub * px / 1e6:.1f} Mpx), "
                    "energy score, reference eval_ood_measure flow on the CPU",
        "cpu": {"metric_s": t_m8, "mpix_s": sub * px / t_m8 / 1e6, "result_hex": hexes(ref8)},
        "gpu": {"metric_ms_device_inputs": g_m8_ms, "mpix_s_device_inputs": sub * px / g_m8_ms / 1e3,
                "metric_e2e_ms_host_numpy_inputs": g_m8_e2e_ms, "metric_e2e_mpix_s": sub * px / g_m8_e2e_ms / 1e3,
                "result_hex": hexes(got8)},
        "metric_equal": hexes(got8) == hexes(ref8) == hexes(got8h),
        "metric_stage_speedup_e2e": t_m8 * 1e3 / g_m8_e2e_ms,
    }
    return out


def cub_yardstick(sizes):
    """cub::DeviceRadixSort on the same box (tools/cub_yardstick, a standalone binary that is NOT part of the library),
    so the repo's own sort has an external anchor; torch.sort (CUB pairs with int64 indices) when it was not built."""
    import subprocess
    exe = os.path.join(ROOT, "tools", "cub_yardstick")
    if os.path.exists(exe):
        try:
            r = subprocess.run([exe] + [str(n) for n in sizes], capture_output=True, text=True, timeout=120)
            rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
            if rows:
                return {"source": "tools/cub_yardstick (cub::DeviceRadixSort, CUDA 12.9 CUB)", "rows": rows}
        except Exception as e:
            err = repr(e)
    rows = []
    for n in sizes:
        k = torch.randint(-2 ** 31, 2 ** 31 - 1, (n,), device="cuda", dtype=torch.int32)
        torch.sort(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.sort(k)
        e1.record()
        torch.cuda.synchronize()
        rows.append({"n": n, "torch_sort_ms": e0.elapsed_time(e1), "torch_sort_gkeys_s": n / e0.elapsed_time(e1) / 1e6})
    return {"source": "torch.sort (CUB SortPairs with int64 indices: a weaker yardstick)", "rows": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the metrics / M2F side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-sweep", action="store_true", help="skip the cfg-4 evaluation sweep (extra.eval_sweep)")
    ap.add_argument("--sweep-images", type=int, default=2000, help="images of the cfg-4 sweep (BASELINE: 2000)")
    ap.add_argument("--continuous-frames", type=int, default=256, help="frames of the T ~= N metric-stage variant")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    dist, rank, world, local = dist_setup(args.gpus)
    from multishiftseg_b200 import _lib as L, deeplab
    lib = L.load()

    g = torch.Generator(device="cuda").manual_seed(2000 + rank)
    logits = torch.randn((B_PER_GPU, C, H, W), device="cuda", generator=g)
    out = {k: torch.empty((B_PER_GPU, H, W), device="cuda") for k in WHICH}
    mask = sum(L.SCORE_BITS[k] for k in WHICH)
    st = torch.cuda.current_stream().cuda_stream

    def step():
        rc = lib.mss_deeplab_score(logits.data_ptr(), B_PER_GPU, C, H * W, mask, out["energy"].data_ptr(),
                                   out["maxlogit"].data_ptr(), 0, out["entropy"].data_ptr(), 0, 0, 0, 1, 0, None, st)
        if rc:
            raise RuntimeError(L.last_error())

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    l0 = L.launch_count()
    ms = timed(step, args.steps, dist, sampler)
    launches = L.launch_count() - l0
    px_per_step = world * B_PER_GPU * H * W
    value = px_per_step * args.steps / ms / 1e3                   # Mpix/s, whole job
    peak, peak_src, peak_doc = peaks()
    kernel_ms = ms / args.steps                                    # one launch per step
    achieved = B_PER_GPU * H * W * BYTES_PER_PX / kernel_ms / 1e6  # GB/s per GPU
    if dist is not None:
        t = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        launches = int(t.item())

    # ---- e2e: host buffers through the C-ABI host entry point -------------------------------------------
    e2e_steps = max(3, min(args.steps, 10))
    h_logits = torch.empty((B_PER_GPU, C, H, W), dtype=torch.float32, pin_memory=True)
    h_logits.copy_(logits)
    h_out = {k: torch.empty((B_PER_GPU, H, W), dtype=torch.float32, pin_memory=True) for k in WHICH}
    nbytes = lib.mss_deeplab_score_host_scratch_bytes(B_PER_GPU, C, H * W, mask)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")

    def e2e_step():
        deeplab.score_maps_host(h_logits, WHICH, scratch=scratch, out=h_out)

    for _ in range(3):
        e2e_step()
    e2e_ms = timed(e2e_step, e2e_steps, dist, sampler)
    e2e_value = px_per_step * e2e_steps / e2e_ms / 1e3
    torch.cuda.synchronize()
    same = all(torch.equal(h_out[k], out[k].cpu()) for k in WHICH)
    del scratch

    # the box's own ceiling for that step: the SAME bytes (2.55 GB in, 0.40 GB out per GPU) as plain pinned copies on two
    # streams, no kernel, all ranks at once -- what the host fabric gives N concurrent GPUs
    side = torch.cuda.Stream()
    d_in = logits
    d_out = [out[k] for k in WHICH]

    def copy_step():
        d_in.copy_(h_logits, non_blocking=True)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for k, t in zip(WHICH, d_out):
                h_out[k].copy_(t, non_blocking=True)
        torch.cuda.current_stream().wait_stream(side)

    def h2d_step():
        d_in.copy_(h_logits, non_blocking=True)

    for _ in range(2):
        copy_step()
    copy_ms = timed(copy_step, 5, dist) / 5
    h2d_ms = timed(h2d_step, 5, dist) / 5
    h2d_bytes = B_PER_GPU * C * H * W * 4
    del h_logits, h_out, d_in, d_out

    line = {
        "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "deeplab_score_vec4_kernel<19,true,false>",
                     "algorithmic_bytes_per_launch": B_PER_GPU * H * W * BYTES_PER_PX, "peak_source": peak_src},
        "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": B_PER_GPU * H * W * 4 * len(WHICH), "steps": e2e_steps,
                "ms_per_step": e2e_ms / e2e_steps, "matches_device_path": bool(same),
                "api": "multishiftseg_b200.deeplab.score_maps_host -> mss_deeplab_score_host (pinned host in/out)",
                # all ranks copying at once (max over ranks): the ceiling the host side of this box gives the step
                "h2d_ceiling_GBs_per_gpu": h2d_bytes / h2d_ms / 1e6,
                "copy_only_ms_per_step": copy_ms,
                "copy_only_ceiling_value": px_per_step / copy_ms / 1e3,
                "frac_of_copy_ceiling": (e2e_value * copy_ms * 1e3) / px_per_step},
        "gpu_launches": launches,
    }
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            line["roofline"]["traffic"] = json.load(open(tr)).get("deeplab_score_vec4_kernel")
        except Exception:
            pass
    del logits, out
    torch.cuda.empty_cache()
    extra = {}

    # ---- the real multi-GPU path, at every N: cfg-4 sweep (key-range exchange + integer prefix merge) ------------
    if not args.no_sweep:
        import bench_sweep
        env = bench_sweep.Env(dist, rank, world, local)
        try:
            # N > 1: exchange="stream" -- every batch is partitioned by owner right behind its scoring kernel and moved
            # over NVLink by the copy engines while the next batch is scored (MSS_BENCH_EXCHANGE overrides, for A/B runs)
            sweep = bench_sweep.run_cfg4(env, args.sweep_images, steps=3, warmup=1, oracle_check=True,
                                         exchange=os.environ.get("MSS_BENCH_EXCHANGE", "stream") if world > 1 else "auto")
            cont = bench_sweep.run_continuous(env, args.continuous_frames, steps=3, warmup=1)
            if rank == 0:
                extra["eval_sweep"] = sweep
                extra["eval_sweep"]["continuous"] = cont
        except Exception as e:   # side measurements must never take the headline down
            if rank == 0:
                extra["eval_sweep"] = {"error": repr(e)}

    if rank == 0 and world == 1 and not args.no_extra:
        try:
            extra.update({"metrics": extra_metrics_stage(), "m2f": extra_m2f(), "m2f_chain": extra_m2f_chain(),
                          "confusion": extra_confusion(),
                          "head": extra_head(), "mask_gemm": extra_mask_gemm(), "backward": extra_backward()})
        except Exception as e:
            extra["error"] = repr(e)
    if sampler is not None:
        line["clocks"] = sampler.summary()
    if rank == 0 and world == 1 and not args.no_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        mpix, reps, el = cpu_reference_scoring(sample_images=2)
        line["cpu_baseline"] = {"value": mpix, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{reps} passes over 2 of the 16 images (1024x2048) in {el:.1f} s, torch-CPU "
                                          "restatement of deepv3.py:251-253 + max-logit + entropy (oracle/scoring_oracle.py)"}
        try:
            extra["cpu_legs"] = cpu_legs()
        except Exception as e:
            extra["cpu_legs"] = {"error": repr(e)}
    if extra:
        line["extra"] = extra
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
