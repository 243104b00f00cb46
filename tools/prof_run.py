"""Short driver for ncu: one pass of each hot kernel at bench sizes (not a benchmark).
argv[1]: which part to run (score|eval|m2f|gemm|all), argv[2]: eval images"""
import sys, torch
sys.path.insert(0, ".")
from multishiftseg_b200 import deeplab, m2f, metric
what = sys.argv[1] if len(sys.argv) > 1 else "all"
imgs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
H, W = 1024, 2048
g = torch.Generator(device="cuda").manual_seed(0)
if what in ("score", "all"):
    x = torch.randn((16, 19, H, W), device="cuda", generator=g)   # the bench launch: 16 images
    for _ in range(2):
        out = deeplab.score_maps(x, ("maxlogit", "energy", "entropy"))
    del x
if what in ("eval", "all"):
    n = imgs * H * W
    lab = torch.zeros(n, dtype=torch.uint8, device="cuda")
    r = torch.rand(n, device="cuda", generator=g)
    lab[r < 0.05] = 1
    lab[r > 0.95] = 255
    s = torch.randn(n, device="cuda", generator=g) + (lab == 1) * 1.5
    del r
    for _ in range(2):
        print(metric.eval_ood_measure(s, lab))
if what in ("m2f", "all"):
    cls = 3.0 * torch.randn((2, 100, 20), device="cuda", generator=g)
    lo = 4.0 * torch.randn((2, 100, 256, 512), device="cuda", generator=g)
    for _ in range(2):
        m2f.anomaly_score_from_lowres(cls, lo, (H, W), (H, W))
        m2f.post_head_inference(cls, lo, (H, W), extra_channels=False)
if what in ("gemm", "all"):
    feat = torch.relu(torch.randn((8, 256, 512, 1024), device="cuda", generator=g))       # DeepLab head features
    wc = torch.randn((19, 256), device="cuda", generator=g) / 16
    wo = torch.randn((19, 256), device="cuda", generator=g) / 16
    for _ in range(2):
        deeplab.head_scores(feat, wc, wo)
    del feat
    mf = torch.randn((8, 256, 256, 512), device="cuda", generator=g)                      # Mask2Former mask features
    emb = torch.randn((8, 100, 256), device="cuda", generator=g) / 16
    for _ in range(2):
        m2f.mask_logits(emb, mf)
torch.cuda.synchronize()
