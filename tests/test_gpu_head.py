"""GPU parity of the fused DeepLab head (deepv3.py:279-283): two 1x1 convolutions + energy on tcgen05
(3xTF32) vs torch's fp32 conv2d / logsumexp on the CPU.

Tolerance: rtol 1e-5 on the energy / upsampled anomaly score (north_star); for the raw logits, which are sums of
256 products that cancel, rtol 1e-5 + atol 1e-5 * sum_k |f_k||w_k| (the fp32 accumulation-order spread of the
CPU reference itself is of that size)."""
import numpy as np
import pytest
import torch

from oracle import scoring_oracle as so

pytestmark = pytest.mark.gpu


def make(B, K, h, w, C=19, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    feat = torch.relu(scale * torch.randn((B, K, h, w), generator=g))          # post-ReLU decoder features
    w_cls = torch.randn((C, K, 1, 1), generator=g) / K ** 0.5
    w_ood = torch.randn((C, K, 1, 1), generator=g) / K ** 0.5
    return feat, w_cls, w_ood


def check(feat, w_cls, w_ood, size=None):
    from multishiftseg_b200 import deeplab
    dec1, score, dec2 = deeplab.head_scores(feat.cuda(), w_cls.cuda(), w_ood.cuda(), size=size, want_dec2=True)
    r1, rs, r2 = so.deeplab_head(feat, w_cls, w_ood, size=size)
    C, K = w_cls.shape[0], w_cls.shape[1]
    for got, want, wt in ((dec1, r1, w_cls), (dec2, r2, w_ood)):
        mag = torch.einsum("bkhw,ck->bchw", feat.abs(), wt.reshape(C, K).abs())
        err = (got.cpu() - want).abs()
        assert bool((err <= 1e-5 * want.abs() + 1e-5 * mag).all()), float((err / (mag + 1e-30)).max())
    np.testing.assert_allclose(score.cpu().numpy(), rs.numpy(), rtol=1e-5, atol=2e-6)
    return dec1, score


@pytest.mark.parametrize("B,K,h,w", [(1, 256, 16, 64), (2, 256, 37, 41), (1, 64, 8, 16), (3, 32, 5, 7), (1, 256, 128, 256),
                                     (2, 96, 9, 33), (1, 160, 20, 20), (1, 192, 3, 130), (2, 224, 17, 19), (150, 32, 4, 40)])
def test_head_vs_torch_fp32(B, K, h, w):
    check(*make(B, K, h, w, seed=B * 1000 + K + h))


def test_model_shape_with_upsample():
    """cfg-5-like: 1080 x 1920 frame, head at 540 x 960, anomaly score upsampled with align_corners=True."""
    feat, w_cls, w_ood = make(1, 256, 540, 960, seed=5)
    check(feat, w_cls, w_ood, size=(1080, 1920))


def test_fewer_classes_and_large_values():
    feat, w_cls, w_ood = make(1, 128, 24, 40, C=7, seed=3, scale=30.0)
    check(feat, w_cls, w_ood)


def test_only_energy_requested_and_consistency_with_two_step_path():
    """head_scores == conv (any exact method) -> deeplab.anomaly_score, i.e. the fused call changes no result."""
    from multishiftseg_b200 import deeplab
    feat, w_cls, w_ood = make(2, 256, 64, 128, seed=9)
    dec1, score, dec2 = deeplab.head_scores(feat.cuda(), w_cls.cuda(), w_ood.cuda(), size=(128, 256), want_dec2=True)
    two_step = deeplab.anomaly_score(dec2, (128, 256))
    np.testing.assert_allclose(score.cpu().numpy(), two_step.cpu().numpy(), rtol=1e-6, atol=1e-6)


def test_unsupported_shapes_raise():
    from multishiftseg_b200 import _lib as L, deeplab
    feat, w_cls, w_ood = make(1, 48, 4, 4)                  # K not a multiple of 32
    with pytest.raises(L.MssError):
        deeplab.head_scores(feat.cuda(), w_cls.cuda(), w_ood.cuda())
    with pytest.raises(L.MssError):
        deeplab.head_scores(feat, w_cls, w_ood)             # CPU tensors: no fallback
