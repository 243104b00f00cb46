"""Timing of the scoring-path backward kernels."""
import sys, torch
sys.path.insert(0, ".")
from multishiftseg_b200 import deeplab
def ev(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
g = torch.Generator(device="cuda").manual_seed(0)
B, C, H, W = 16, 19, 1024, 2048
x = torch.randn((B, C, H, W), device="cuda", generator=g).requires_grad_(True)
go = torch.randn((B, H, W), device="cuda", generator=g)
s = deeplab.energy_func(x)
t = ev(lambda: torch.autograd.grad(s, x, go, retain_graph=True))
px = B * H * W
print(f"energy backward 16x19x1024x2048: {t:7.3f} ms  {px*(2*C*4+4)/t/1e6:7.0f} GB/s algorithmic")
xt = x.detach().clone().requires_grad_(True)
st = -torch.logsumexp(xt, dim=1)
t2 = ev(lambda: torch.autograd.grad(st, xt, go, retain_graph=True))
print(f"torch autograd of -logsumexp:     {t2:7.3f} ms")
del x, xt, s, st
h, w = 512, 1024
x = torch.randn((B, C, h, w), device="cuda", generator=g).requires_grad_(True)
s = deeplab.anomaly_score(x, (H, W))
t = ev(lambda: torch.autograd.grad(s, x, go, retain_graph=True))
print(f"anomaly_score backward (512x1024 head -> 1024x2048): {t:7.3f} ms  {B*(h*w*2*C*4+H*W*4)/t/1e6:7.0f} GB/s algorithmic")
xt = x.detach().clone().requires_grad_(True)
st = torch.nn.functional.interpolate((-torch.logsumexp(xt, dim=1)).unsqueeze(1), size=(H, W), mode="bilinear", align_corners=True).squeeze(1)
t2 = ev(lambda: torch.autograd.grad(st, xt, go, retain_graph=True))
print(f"torch autograd of the same:        {t2:7.3f} ms")
