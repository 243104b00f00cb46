"""GPU: streaming evaluator (single GPU) and, when the box has >= 2 GPUs, the NCCL multi-GPU path."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import gen_inputs as gi
from oracle import c_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_streaming_equals_one_shot_and_oracle():
    from multishiftseg_b200 import metric
    from multishiftseg_b200.evaluator import StreamingEvaluator
    s, l = gi.metric_case(31, 1_500_000, "cont", label_dtype="int64")
    ev = StreamingEvaluator(s.size, distributed=False)
    for a in range(0, s.size, 250_000):          # six "batches", as the tester loop would feed them
        ev.update(torch.from_numpy(s[a:a + 250_000]).cuda(), torch.from_numpy(l[a:a + 250_000]).cuda())
    got = tuple(float(v) for v in ev.compute())
    assert got == c_oracle.eval_ood_measure(s, l)
    assert got == tuple(float(v) for v in metric.eval_ood_measure(s, l))
    ev.reset()
    ev.update(np.array([.1, .4, .35, .8], np.float32), np.array([0, 0, 1, 1]))
    assert tuple(float(v) for v in ev.compute()) == (0.75, float.fromhex("0x1.aaaaaaaaaaaaap-1"), 0.5)


def test_capacity_overflow_is_an_error():
    from multishiftseg_b200 import _lib
    from multishiftseg_b200.evaluator import StreamingEvaluator
    ev = StreamingEvaluator(100, distributed=False)
    ev.update(torch.zeros(1000, device="cuda"), torch.zeros(1000, dtype=torch.uint8, device="cuda"))
    with pytest.raises(_lib.MssError):
        ev.compute()


def test_multi_gpu_bit_exact():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    # the same through the C entry point a C / C++ binder would call (its own ncclComm_t)
    cmd[-1] = os.path.join(ROOT, "tests", "dist_c_abi_check.py")
    cmd[cmd.index("29533")] = "29534"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
