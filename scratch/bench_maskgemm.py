"""Timing of the Mask2Former mask-logit GEMM (8 x [100 x 256] x [256 x 256 x 512] -> decoder-resolution masks)."""
import sys, torch
sys.path.insert(0, ".")
from multishiftseg_b200 import m2f
B, Q, K, h, w = 8, 100, 256, 256, 512
g = torch.Generator(device="cuda").manual_seed(0)
feat = torch.randn((B, K, h, w), device="cuda", generator=g)
emb = torch.randn((B, Q, K), device="cuda", generator=g) / 16
cls = 3.0 * torch.randn((B, Q, 20), device="cuda", generator=g)
def ev(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t = ev(lambda: m2f.mask_logits(emb, feat))
px = B * h * w
byt = px * (K * 4 + Q * 4)
print(f"mask GEMM (tcgen05 3xTF32): {t:7.3f} ms  {byt/t/1e6:7.0f} GB/s algorithmic  ({t*1e3/B:6.1f} us/image)  "
      f"{2*px*Q*K/t/1e9:6.1f} TFLOP/s fp32-equivalent")
torch.backends.cuda.matmul.allow_tf32 = False
t2 = ev(lambda: torch.einsum("bqc,bchw->bqhw", emb, feat))
print(f"torch einsum fp32 (cuBLAS SGEMM):       {t2:7.3f} ms  ({t2*1e3/B:6.1f} us/image)")
torch.backends.cuda.matmul.allow_tf32 = True
t3 = ev(lambda: torch.einsum("bqc,bchw->bqhw", emb, feat))
print(f"torch einsum tf32 allowed (1xTF32):     {t3:7.3f} ms  ({t3*1e3/B:6.1f} us/image)")
torch.backends.cuda.matmul.allow_tf32 = False
t4 = ev(lambda: m2f.anomaly_score_from_features(cls, emb, feat, (4 * h, 4 * w), (4 * h, 4 * w)))
print(f"GEMM + fused scoring (features -> anomaly score 1024x2048): {t4:7.3f} ms  ({t4*1e3/B:6.1f} us/image)")
