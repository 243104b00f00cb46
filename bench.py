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
  extra      the other two kernels of the path (exact metrics, Mask2Former fused inference), each with
             its own throughput / roofline fraction, so one run documents the whole path

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
    """`--impl reference`: the reference's own CPU path (oracle port) for the same metric / config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import scoring_oracle as so
    torch.set_num_threads(os.cpu_count() or 1)
    sample = 2                                   # images per step (of the 16-image batch): bounded sample
    g = torch.Generator().manual_seed(0)
    x = torch.randn((sample, C, H, W), generator=g)

    def step():
        return so.maxlogit_score(x), so.energy_func(x), so.entropy_score(x)

    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    mpix = args.steps * sample * H * W / el / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample} of 16 images per step, CPU"},
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} steps x {sample} images (1024x2048) of the 16-image batch, "
                                   "torch-CPU restatement of deepv3.py:251-253 + max-logit + entropy"},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
def extra_metrics_stage(images=4):
    """Exact AUROC/AP/FPR95 on `images` x 1024 x 2048 score/label maps resident in HBM."""
    from multishiftseg_b200 import metric
    n = images * H * W
    g = torch.Generator(device="cuda").manual_seed(4000)
    lab = torch.zeros(n, dtype=torch.uint8, device="cuda")
    r = torch.rand(n, device="cuda", generator=g)
    lab[r < 0.05] = 1
    lab[r > 0.95] = 255
    s = torch.randn(n, device="cuda", generator=g) + (lab == 1) * 1.5
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
    # isolate the sort (dominant kernel family): pairs already built
    buf = metric.PairBuffer(n, "cuda")
    buf.append(s, lab)
    m = buf.read_state()[0]
    keys0, labs0 = buf.keys.clone(), buf.labs.clone()
    sort_ms = []
    for _ in range(5):
        buf.keys.copy_(keys0)
        buf.labs.copy_(labs0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        metric.sort_pairs(buf.keys, buf.labs, m)
        e1.record()
        torch.cuda.synchronize()
        sort_ms.append(e0.elapsed_time(e1))
    sm = statistics.median(sort_ms)
    peak, _, _ = peaks()
    return {"workload": f"exact AUROC/AP/FPR95, {images}x1024x2048 px in HBM (90/5/5 % ID/OOD/ignore)",
            "mpix_s": n / t / 1e6, "images_s": images / t, "ms": t * 1e3, "valid_pairs": m,
            "result": [float(x) for x in res],
            "sort": {"ms": sm, "gkeys_s": m / sm / 1e6,
                     "algorithmic_GBs": m * 5 / sm / 1e6,       # 5 B/pair read once (SURVEY 8d lower bound)
                     "implementation_GBs": m * 44 / sm / 1e6,   # 4 + 4 x (5 + 5) B/pair actually moved
                     "frac_of_hbm_peak_impl": m * 44 / sm / 1e6 / peak}}


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
    out["workload"] = f"cfg3: M2F semantic_inference Q=100 C=19+1 256x512->1024x2048 batch {batch}"
    out["fma_roofline_TFLOPs"] = 148 * 128 * 2 * 1.965e9 / 1e12
    # binding pipe of the tcgen05 kernel: the two MUFU ops (ex2 + rcp) of each of the Q sigmoids per pixel,
    # 16 lanes/clk/SM -> 2 * Q * px / (16 * 148 * f_SM)
    floor_ms = 2 * 100 * batch * H * W / (16 * 148 * 1.965e9) * 1e3
    out["mufu_floor_ms"] = floor_ms
    out["frac_of_mufu_roofline"] = floor_ms / out["anomaly_score"]["ms"]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the metrics / M2F side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
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
    del h_logits, scratch

    line = {
        "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B_PER_GPU, "classes": C, "frame": [H, W],
                   "l2": "input 2.55 GB per step >> 126 MB L2, no flush needed",
                   "parallelism": f"images sharded over {world} GPU(s), no data-path collective"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "deeplab_score_vec4_kernel<19,true,false>",
                     "algorithmic_bytes_per_launch": B_PER_GPU * H * W * BYTES_PER_PX, "peak_source": peak_src},
        "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": B_PER_GPU * C * H * W * 4,
                "d2h_bytes_per_step": B_PER_GPU * H * W * 4 * len(WHICH), "steps": e2e_steps,
                "ms_per_step": e2e_ms / e2e_steps, "matches_device_path": bool(same),
                "api": "multishiftseg_b200.deeplab.score_maps_host -> mss_deeplab_score_host (pinned host in/out)"},
        "gpu_launches": launches,
    }
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            line["roofline"]["traffic"] = json.load(open(tr)).get("deeplab_score_vec4_kernel")
        except Exception:
            pass

    if rank == 0 and world == 1 and not args.no_extra:
        try:
            del logits, out
            torch.cuda.empty_cache()
            line["extra"] = {"metrics": extra_metrics_stage(), "m2f": extra_m2f(), "confusion": extra_confusion(),
                             "head": extra_head(), "mask_gemm": extra_mask_gemm(), "backward": extra_backward()}
        except Exception as e:   # side measurements must never take the headline down
            line["extra"] = {"error": repr(e)}
    if sampler is not None:
        line["clocks"] = sampler.summary()
    if rank == 0 and world == 1 and not args.no_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        mpix, reps, el = cpu_reference_scoring(sample_images=2)
        line["cpu_baseline"] = {"value": mpix, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{reps} passes over 2 of the 16 images (1024x2048) in {el:.1f} s, torch-CPU "
                                          "restatement of deepv3.py:251-253 + max-logit + entropy (oracle/scoring_oracle.py)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
