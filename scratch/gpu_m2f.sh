#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_m2f.py -x -q 2>&1 | tail -30
timeout 120 python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from multishiftseg_b200 import m2f
H, W = 1024, 2048
g = torch.Generator(device="cuda").manual_seed(0)
cls = 3.0 * torch.randn((8, 100, 20), device="cuda", generator=g)
lo = 4.0 * torch.randn((8, 100, 256, 512), device="cuda", generator=g)
for flags, name in ((0, "tcgen05"), (4, "mma.sync"), (2, "ffma")):
    for what, fn in (("anomaly", lambda: m2f.anomaly_score_from_lowres(cls, lo, (H, W), (H, W), flags=flags)),
                     ("semseg19", lambda: m2f.post_head_inference(cls, lo, (H, W), extra_channels=False, flags=flags))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:9s} {what:9s} {ms:7.3f} ms/batch8  {ms/8*1e3:7.1f} us/img  {8*H*W/ms/1e6:6.2f} Gpix/s")
PY
