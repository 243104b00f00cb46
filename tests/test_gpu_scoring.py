"""GPU parity, scoring stage: CUDA kernels vs the outputs of the reference's own function bodies
(tests/golden/scoring_golden.npz) and vs the torch-fp32 oracle on seeded inputs.
Tolerance (north_star): 1e-5 relative, plus atol 2e-6 for values that cancel to ~0."""
import os

import numpy as np
import pytest
import torch

import gen_inputs as gi
from oracle import scoring_oracle as so

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 2e-6

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "scoring_golden.npz"))
D = {k: torch.from_numpy(v) for k, v in gi.scoring_inputs().items()}


@pytest.fixture(scope="module")
def dl():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no fallback)"
    from multishiftseg_b200 import deeplab
    return deeplab


def close(got, want, rtol=RTOL, atol=ATOL):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    want = want.detach().cpu().numpy() if isinstance(want, torch.Tensor) else want
    assert got.shape == want.shape, (got.shape, want.shape)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


def test_energy_golden(dl):
    close(dl.energy_func(D["dl_logit"].cuda()), G["dl_energy"])
    close(dl.energy_func(D["dl_logit_odd"].cuda()), G["dl_energy_odd"])      # HW % 4 != 0: generic kernel


def test_upsample_golden(dl):
    close(dl.Upsample(D["up_in"].cuda(), (24, 40)), G["up_x2"])
    close(dl.Upsample(D["up_in"].cuda(), (31, 53)), G["up_odd"])
    close(dl.anomaly_score(D["dl_logit"].cuda(), (48, 80)), G["dl_anomaly_x2"])


@pytest.mark.parametrize("shape", [(1, 19, 64, 128), (3, 19, 33, 52), (2, 19, 1, 4), (2, 7, 16, 16), (1, 32, 9, 5)])
def test_all_scores_vs_oracle(dl, shape):
    g = torch.Generator().manual_seed(sum(shape))
    x = 4.0 * torch.randn(shape, generator=g)
    out = dl.score_maps(x.cuda(), ("energy", "maxlogit", "msp", "entropy"))
    close(out["energy"], so.energy_func(x))
    assert torch.equal(out["maxlogit"].cpu(), so.maxlogit_score(x))          # pure max: bit-exact
    close(out["msp"], so.msp_score(x))
    close(out["entropy"], so.entropy_score(x), atol=5e-6)


def test_subset_of_scores_and_large_logits(dl):
    g = torch.Generator().manual_seed(9)
    x = 40.0 * torch.randn((2, 19, 32, 64), generator=g)      # large magnitudes: stable logsumexp needed
    out = dl.score_maps(x.cuda(), ("maxlogit", "energy"))
    assert set(out) == {"maxlogit", "energy"}
    close(out["energy"], so.energy_func(x))


def test_upsample_align_corners_false_matches_torch(dl):
    g = torch.Generator().manual_seed(2)
    x = torch.randn((2, 5, 17, 23), generator=g)
    want = torch.nn.functional.interpolate(x, size=(68, 92), mode="bilinear", align_corners=False)
    close(dl.Upsample(x.cuda(), (68, 92), align_corners=False), want)


def test_cfg5_shape_half_res_upsample(dl):
    """cfg-5: 540 x 960 head output -> 1080 x 1920 (deepv3.py:283)."""
    g = torch.Generator().manual_seed(3)
    x = 3.0 * torch.randn((1, 19, 540, 960), generator=g)
    close(dl.anomaly_score(x.cuda(), (1080, 1920)), so.deeplab_anomaly_score(x, (1080, 1920)))


def test_fused_evaluator_append_equals_separate_path(dl):
    """score + ignore-mask + key build in one kernel == score map then eval_ood_measure."""
    from multishiftseg_b200 import metric
    g = torch.Generator().manual_seed(4)
    B, H, W = 2, 128, 256
    x = torch.randn((B, 19, H, W), generator=g)
    r = torch.rand((B, H, W), generator=g)
    lab = torch.where(r < 0.05, 1, torch.where(r > 0.95, 255, 0)).to(torch.int64)
    x = torch.where((lab == 1).unsqueeze(1), 0.5 * x, 2.0 * x)
    for ldt in (torch.int64, torch.uint8):
        buf = metric.PairBuffer(B * H * W, "cuda")
        out = dl.score_maps(x.cuda(), ("energy", "entropy"), labels=lab.to(ldt).cuda(), evaluator=buf, key="energy")
        fused = metric._finish(buf)
        sep = metric.eval_ood_measure(out["energy"], lab.cuda())
        assert tuple(map(float, fused)) == tuple(map(float, sep))
    # and the whole thing equals the CPU reference flow on the oracle's score map within metric exactness
    from oracle import c_oracle
    assert tuple(map(float, sep)) == c_oracle.eval_ood_measure(out["energy"].cpu().numpy(), lab.numpy())


def test_host_buffer_entry_point(dl):
    g = torch.Generator().manual_seed(6)
    x = torch.randn((5, 19, 64, 96), generator=g).pin_memory()
    out = dl.score_maps_host(x, ("energy", "maxlogit", "entropy"))
    ref = dl.score_maps(x.cuda(), ("energy", "maxlogit", "entropy"))
    for k in ref:
        assert torch.equal(out[k], ref[k].cpu()), k


def test_bench_shape_linearity_property(dl):
    """Full cfg-2 plane size (1024 x 2048): energy(x + c) == energy(x) - c up to rounding (shift property)."""
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((2, 19, 1024, 2048), device="cuda", generator=g)
    e0 = dl.energy_func(x)
    e1 = dl.energy_func(x + 3.0)
    assert torch.allclose(e1, e0 - 3.0, rtol=1e-5, atol=1e-5)
    ref = -torch.logsumexp(x, dim=1)
    assert torch.allclose(e0, ref, rtol=RTOL, atol=ATOL)
