"""The remote-append exchange kernel on ONE GPU with every destination a local buffer (development aid): the kernel's own
cost without NVLink, timeable with CUDA events and profilable with ncu.   python tools/exch_local.py [parts] [keys]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multishiftseg_b200 import metric  # noqa: E402
from multishiftseg_b200.evaluator import CudaBackend, choose_splitters  # noqa: E402

parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 28
be = CudaBackend()
g = torch.Generator(device="cuda").manual_seed(3)
buf = metric.PairBuffer(n, "cuda")
buf.reset()
for _ in range(4):
    s = torch.randn(n // 4, device="cuda", generator=g)
    lab = (torch.rand(n // 4, device="cuda", generator=g) < 0.05).to(torch.uint8)
    buf.append(s, lab)
del s, lab
m, n_pos, _, _ = buf.read_state()
neg, pos = buf.streams(m, n_pos)
hist = be.histogram(neg, m - n_pos, 16, 8) + be.histogram(pos, n_pos, 16, 8)
spl = choose_splitters(hist.cpu().numpy(), parts)
cap = m // parts + m // (4 * parts) + (1 << 20)
dst = [metric.PairBuffer(cap, "cuda") for _ in range(parts)]


def run():
    for d in dst:
        d.reset()
    be.exchange_append(buf, m - n_pos, n_pos, spl, parts, [d.keys.data_ptr() for d in dst], [d.state.data_ptr() for d in dst], cap)


run()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    for d in dst:
        d.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    be.exchange_append(buf, m - n_pos, n_pos, spl, parts, [d.keys.data_ptr() for d in dst], [d.state.data_ptr() for d in dst], cap)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
got = sum(d.read_state()[0] for d in dst)
print(f"exchange_append local, {parts} destinations, {m} keys: {sorted(ts)[2]:.3f} ms = {m / sorted(ts)[2] / 1e6:.1f} Gkeys/s "
      f"({m * 8 / sorted(ts)[2] / 1e6:.0f} GB/s read+write), delivered {got} ({'ok' if got == m else 'MISMATCH'})")
