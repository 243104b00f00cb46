import sys, torch
sys.path.insert(0, ".")
from multishiftseg_b200 import m2f
from oracle import scoring_oracle as so
h, w, Q, B = [int(v) for v in sys.argv[1:5]]
g = torch.Generator().manual_seed(0)
cls = 3.0 * torch.randn((B, Q, 20), generator=g)
lo = 4.0 * torch.randn((B, Q, h, w), generator=g)
a = m2f.anomaly_score_from_lowres(cls.cuda(), lo.cuda(), (4*h, 4*w), (4*h, 4*w))
torch.cuda.synchronize()
want = so.m2f_anomaly_from_lowres(cls, lo, (4*h, 4*w), (4*h, 4*w))
print("OK", h, w, Q, B, float((a.cpu()-want).abs().max()))
