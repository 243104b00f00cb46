"""Does the achieved bandwidth of the two GEMM kernels depend on the plane stride (power of two or not)?"""
import sys, torch
sys.path.insert(0, ".")
from multishiftseg_b200 import m2f, deeplab
def ev(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
g = torch.Generator(device="cuda").manual_seed(0)
for (h, w) in [(256, 512), (272, 480), (256, 520), (260, 512), (257, 511)]:
    B, Q, K = 8, 100, 256
    feat = torch.randn((B, K, h, w), device="cuda", generator=g)
    emb = torch.randn((B, Q, K), device="cuda", generator=g) / 16
    t = ev(lambda: m2f.mask_logits(emb, feat))
    print(f"mask GEMM {h}x{w}: {t:7.3f} ms  {B*h*w*(K*4+Q*4)/t/1e6:7.0f} GB/s")
    del feat
for (h, w) in [(512, 1024), (540, 960), (512, 1040), (520, 1024)]:
    B, K = 8, 256
    feat = torch.relu(torch.randn((B, K, h, w), device="cuda", generator=g))
    wc = torch.randn((19, K), device="cuda", generator=g) / 16
    wo = torch.randn((19, K), device="cuda", generator=g) / 16
    t = ev(lambda: deeplab.head_scores(feat, wc, wo))
    print(f"head {h}x{w}: {t:7.3f} ms  {B*h*w*(K*4+19*4+4)/t/1e6:7.0f} GB/s")
    del feat
