"""flags: 0 = default fast path (TMA patch + tcgen05/TMEM 3xTF32 contraction, shared-tap 4x1 pixel blocks),
8 = tcgen05 with one pixel per thread, 4 = mma.sync contraction, 2 = FP32-FMA contraction, 1 = generic any-resize kernel.

GPU parity, Mask2Former post-head path: fused kernels vs outputs of the reference's own
semantic_inference / get_anomaly_score bodies (tests/golden/scoring_golden.npz) and the torch oracle.
Tolerance: 1e-5 relative + atol 2e-6 (1 - max and sums of ~100 products cancel toward 0)."""
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
def m2f():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no fallback)"
    from multishiftseg_b200 import m2f
    return m2f


def close(got, want, rtol=RTOL, atol=ATOL):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    want = want.detach().cpu().numpy() if isinstance(want, torch.Tensor) else want
    assert got.shape == want.shape, (got.shape, want.shape)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


@pytest.mark.parametrize("flags", [0, 8, 4, 2, 1], ids=["tma_tcgen05", "tma_tcgen05_pixel", "tma_mmasync", "tma_ffma", "generic"])
def test_fused_from_lowres_golden(m2f, flags):
    cls, lo = D["m2f_cls"].cuda(), D["m2f_mask_lo"].cuda()
    outs = m2f.post_head_inference(cls, lo, (32, 64), flags=flags)
    for b in range(2):
        close(outs[b], G[f"m2f_semseg_{b}"])                      # 19 + K channels, K data dependent
    close(m2f.anomaly_score_from_lowres(cls, lo, (32, 64), (30, 61), flags=flags), G["m2f_anomaly"])


def test_dropin_signatures_on_upsampled_masks(m2f):
    """semantic_inference(mask_cls, mask_pred) / get_anomaly_score(outputs, size) exactly as the reference calls them."""
    up = so.upsample_masks(D["m2f_mask_lo"], (32, 64))
    for b in range(2):
        close(m2f.semantic_inference(D["m2f_cls"][b].cuda(), up[b].cuda()), G[f"m2f_semseg_{b}"])
    got = m2f.get_anomaly_score({"pred_logits_ood": D["m2f_cls"].cuda(), "pred_masks_ood": up.cuda()}, (30, 61))
    close(got, G["m2f_anomaly"])


@pytest.mark.parametrize("flags", [0, 8, 4, 2, 1], ids=["tma_tcgen05", "tma_tcgen05_pixel", "tma_mmasync", "tma_ffma", "generic"])
@pytest.mark.parametrize("hw,crop", [((64, 128), (256, 512)), ((68, 120), (270, 480)), ((36, 40), (141, 157))])
def test_random_vs_oracle(m2f, flags, hw, crop):
    g = torch.Generator().manual_seed(hw[0] * 1000 + hw[1])
    B, Q = 2, 100
    cls = 3.0 * torch.randn((B, Q, 20), generator=g)
    lo = 4.0 * torch.randn((B, Q) + hw, generator=g)
    padded = (4 * hw[0], 4 * hw[1])
    want = so.m2f_post_head(cls, lo, padded, crop)
    got = m2f.post_head_inference(cls.cuda(), lo.cuda(), padded, [crop] * B, flags=flags)
    for b in range(B):
        close(got[b], want[b])
    close(m2f.anomaly_score_from_lowres(cls.cuda(), lo.cuda(), padded, crop, flags=flags),
          so.m2f_anomaly_from_lowres(cls, lo, padded, crop))


def test_non_x4_resize_and_small_q(m2f):
    g = torch.Generator().manual_seed(77)
    cls = 2.0 * torch.randn((1, 37, 20), generator=g)
    lo = 3.0 * torch.randn((1, 37, 20, 30), generator=g)
    want = so.m2f_post_head(cls, lo, (50, 77), (50, 77))
    got = m2f.post_head_inference(cls.cuda(), lo.cuda(), (50, 77))
    close(got[0], want[0])
    # Q not a multiple of the TMA chunk on the fast path
    lo4 = 3.0 * torch.randn((2, 37, 16, 32), generator=g)
    cls2 = 2.0 * torch.randn((2, 37, 20), generator=g)
    want = so.m2f_post_head(cls2, lo4, (64, 128), (64, 128))
    got = m2f.post_head_inference(cls2.cuda(), lo4.cuda(), (64, 128))
    for b in range(2):
        close(got[b], want[b])


def test_cfg5_padded_shape(m2f):
    """cfg-5: 1080 x 1920 image padded to 1088 x 1920 (size divisibility 32), masks 272 x 480, crop back."""
    g = torch.Generator().manual_seed(5)
    cls = 3.0 * torch.randn((1, 100, 20), generator=g)
    lo = 4.0 * torch.randn((1, 100, 272, 480), generator=g)
    got = m2f.anomaly_score_from_lowres(cls.cuda(), lo.cuda(), (1088, 1920), (1080, 1920))
    want = so.m2f_anomaly_from_lowres(cls, lo, (1088, 1920), (1080, 1920))
    close(got, want)


def test_full_resolution_tile_consistency(m2f):
    """cfg-3 frame size: the TMA-tiled kernel and the generic kernel agree on a 1024 x 2048 frame."""
    g = torch.Generator(device="cuda").manual_seed(8)
    cls = 3.0 * torch.randn((1, 100, 20), device="cuda", generator=g)
    lo = 4.0 * torch.randn((1, 100, 256, 512), device="cuda", generator=g)
    a = m2f.anomaly_score_from_lowres(cls, lo, (1024, 2048), (1024, 2048), flags=0)
    b = m2f.anomaly_score_from_lowres(cls, lo, (1024, 2048), (1024, 2048), flags=1)
    c = m2f.anomaly_score_from_lowres(cls, lo, (1024, 2048), (1024, 2048), flags=2)
    d = m2f.anomaly_score_from_lowres(cls, lo, (1024, 2048), (1024, 2048), flags=4)
    e = m2f.anomaly_score_from_lowres(cls, lo, (1024, 2048), (1024, 2048), flags=8)
    assert torch.allclose(e, b, rtol=RTOL, atol=ATOL)
    assert torch.allclose(a, b, rtol=RTOL, atol=ATOL)
    assert torch.allclose(c, b, rtol=RTOL, atol=ATOL)
    assert torch.allclose(d, b, rtol=RTOL, atol=ATOL)
    s = m2f.post_head_inference(cls, lo, (1024, 2048), extra_channels=False)[0]
    assert torch.allclose(1 - s.max(0)[0], a, rtol=RTOL, atol=ATOL)
    # probabilities: every channel in [0, sum_q P] and the 19-channel sum <= Q
    assert float(s.min()) >= 0.0 and float(s.sum(0).max()) <= 100.0 + 1e-3
