"""cfg-3 timing of the M2F fused kernel variants (flags: 0 default, 8 tcgen05 pixel-per-thread, 4 mma.sync, 2 ffma)."""
import sys, torch
sys.path.insert(0, ".")
from multishiftseg_b200 import m2f
B = 8
g = torch.Generator(device="cuda").manual_seed(0)
cls = 3.0 * torch.randn((B, 100, 20), device="cuda", generator=g)
lo = 4.0 * torch.randn((B, 100, 256, 512), device="cuda", generator=g)
def ev(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for flags in [int(a) for a in sys.argv[1:]] or [0, 8]:
    ta = ev(lambda: m2f.anomaly_score_from_lowres(cls, lo, (1024, 2048), (1024, 2048), flags=flags))
    ts = ev(lambda: m2f.post_head_inference(cls, lo, (1024, 2048), extra_channels=False, flags=flags))
    print(f"flags={flags}: anomaly {ta*1e3/B:7.1f} us/img ({B*2.097152/ta:7.2f} Gpix/s)   semseg19 {ts*1e3/B:7.1f} us/img")
