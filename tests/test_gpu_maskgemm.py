"""GPU parity of the Mask2Former mask-logit GEMM (mask2former_transformer_decoder.py:529 / :549):
``einsum("bqc,bchw->bqhw", mask_embed, mask_features)`` on tcgen05 (3xTF32) vs torch's fp32 einsum on the CPU.

Tolerance: the mask logits are sums of 256 products that cancel, so (as for the DeepLab head logits)
rtol 1e-5 + atol 1e-5 * sum_k |e_k||f_k| -- the fp32 accumulation-order spread of the CPU reference itself is of
that size; a float64 reference bounds both.  The score that comes out of the full chain (GEMM -> fused
upsample / sigmoid / contraction / 1 - max) is held to the north_star tolerance rtol 1e-5 (+ atol 2e-6)."""
import numpy as np
import pytest
import torch

from oracle import scoring_oracle as so

pytestmark = pytest.mark.gpu


def make(B, Q, K, h, w, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    feat = scale * torch.randn((B, K, h, w), generator=g)
    embed = torch.randn((B, Q, K), generator=g) / K ** 0.5
    return embed, feat


def check(embed, feat):
    from multishiftseg_b200 import m2f
    got = m2f.mask_logits(embed.cuda(), feat.cuda()).cpu()
    want = so.m2f_mask_logits(embed, feat)
    assert got.shape == want.shape
    mag = torch.einsum("bqc,bchw->bqhw", embed.abs(), feat.abs())
    err = (got - want).abs()
    assert bool((err <= 1e-5 * want.abs() + 1e-5 * mag).all()), float((err / (mag + 1e-30)).max())
    # against float64: the 3xTF32 result must be as close to the exact value as fp32 accumulation is
    exact = torch.einsum("bqc,bchw->bqhw", embed.double(), feat.double())
    e_gpu = float(((got.double() - exact).abs() / (mag.double() + 1e-30)).max())
    e_cpu = float(((want.double() - exact).abs() / (mag.double() + 1e-30)).max())
    assert e_gpu <= max(4 * e_cpu, 2e-6), (e_gpu, e_cpu)
    return got


@pytest.mark.parametrize("B,Q,K,h,w", [(1, 100, 256, 16, 64), (2, 100, 256, 37, 41), (1, 100, 64, 8, 16), (3, 7, 32, 5, 7),
                                       (1, 112, 256, 24, 40), (1, 100, 256, 128, 256), (2, 1, 96, 6, 50), (1, 100, 160, 19, 23),
                                       (1, 57, 192, 16, 16), (2, 100, 224, 11, 40)])
def test_mask_logits_vs_torch_fp32(B, Q, K, h, w):
    check(*make(B, Q, K, h, w, seed=B * 1000 + K + h))


def test_per_image_tables_many_images():
    """More images than SMs: whole images round-robin over the CTAs, the embedding table is reloaded per image."""
    check(*make(151, 100, 64, 9, 31, seed=11))


def test_few_tiles_per_image():
    """Fewer tiles than slices / one group idle: h*w = 100 px (one partial tile), and 3 tiles."""
    check(*make(2, 100, 256, 10, 10, seed=12))
    check(*make(1, 100, 256, 3, 128, seed=13))


def test_large_values():
    check(*make(1, 100, 128, 24, 40, seed=3, scale=30.0))


def test_cfg5_shape_chain_to_anomaly_score():
    """cfg-5: 1080 x 1920 image padded to 1088 x 1920, mask features 272 x 480; GEMM -> fused scoring kernel."""
    from multishiftseg_b200 import m2f
    g = torch.Generator().manual_seed(5)
    embed, feat = make(1, 100, 256, 272, 480, seed=5, scale=2.0)
    cls = 3.0 * torch.randn((1, 100, 20), generator=g)
    got = m2f.anomaly_score_from_features(cls.cuda(), embed.cuda(), feat.cuda(), (1088, 1920), (1080, 1920))
    lo = so.m2f_mask_logits(embed, feat)
    want = so.m2f_anomaly_from_lowres(cls, lo, (1088, 1920), (1080, 1920))
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=2e-6)


def test_unsupported_shapes_raise():
    from multishiftseg_b200 import _lib as L, m2f
    embed, feat = make(1, 100, 48, 4, 4)                    # K not a multiple of 32
    with pytest.raises(L.MssError):
        m2f.mask_logits(embed.cuda(), feat.cuda())
    embed, feat = make(1, 120, 64, 4, 4)                    # Q > 112
    with pytest.raises(L.MssError):
        m2f.mask_logits(embed.cuda(), feat.cuda())
    with pytest.raises(L.MssError):
        m2f.mask_logits(embed, feat)                        # CPU tensors: no fallback
