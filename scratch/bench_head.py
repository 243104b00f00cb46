"""Timing of the fused DeepLab head (8 x 256 x 512 x 1024 features -> dec1 + energy)."""
import sys, torch
sys.path.insert(0, ".")
from multishiftseg_b200 import deeplab
B, K, h, w = 8, 256, 512, 1024
g = torch.Generator(device="cuda").manual_seed(0)
feat = torch.relu(torch.randn((B, K, h, w), device="cuda", generator=g))
wc = torch.randn((19, K), device="cuda", generator=g) / 16
wo = torch.randn((19, K), device="cuda", generator=g) / 16
def ev(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t = ev(lambda: deeplab.head_scores(feat, wc, wo))
px = B * h * w
byt = px * (K * 4 + 19 * 4 + 4)
print(f"fused head: {t:7.3f} ms  {px/t/1e6:7.2f} Gpix(head-res)/s  {byt/t/1e6:7.0f} GB/s algorithmic  ({t*1e3/B:6.1f} us/image)")
import torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = True
t2 = ev(lambda: (F.conv2d(feat, wc.view(19, K, 1, 1)), -torch.logsumexp(F.conv2d(feat, wo.view(19, K, 1, 1)), 1)))
print(f"torch cuDNN (tf32 allowed) two convs + logsumexp: {t2:7.3f} ms")
