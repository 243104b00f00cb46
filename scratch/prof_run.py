"""Short driver for ncu: one pass of each hot kernel at bench sizes (not a benchmark)."""
import sys, torch
sys.path.insert(0, ".")
from multishiftseg_b200 import deeplab, m2f, metric
H, W = 1024, 2048
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((8, 19, H, W), device="cuda", generator=g)
for _ in range(2):
    out = deeplab.score_maps(x, ("maxlogit", "energy", "entropy"))
n = 4 * H * W
lab = torch.zeros(n, dtype=torch.uint8, device="cuda")
r = torch.rand(n, device="cuda", generator=g)
lab[r < 0.05] = 1
lab[r > 0.95] = 255
s = torch.randn(n, device="cuda", generator=g) + (lab == 1) * 1.5
for _ in range(2):
    print(metric.eval_ood_measure(s, lab))
cls = 3.0 * torch.randn((2, 100, 20), device="cuda", generator=g)
lo = 4.0 * torch.randn((2, 100, 256, 512), device="cuda", generator=g)
for _ in range(2):
    m2f.anomaly_score_from_lowres(cls, lo, (H, W), (H, W))
    m2f.post_head_inference(cls, lo, (H, W), extra_channels=False)
torch.cuda.synchronize()
