"""GPU parity of the scoring path's backward (DeepLab half): the trainer differentiates through
energy_func + Upsample (train_deeplab.py:197-198; deepv3.py:251-253, :283; mynn.py:28-33).  Reference = torch autograd
of the oracle's fp32 expressions on the CPU.

Tolerance: rtol 1e-5 (north_star) + atol 1e-6 * max|grad|: a softmax probability near 0 times the incoming gradient has
no relative accuracy in either implementation, and the gather-form adjoint of the bilinear upsample adds its (at most
~9) terms in a different order than ATen's scatter-add."""
import numpy as np
import pytest
import torch

from oracle import scoring_oracle as so

pytestmark = pytest.mark.gpu


def close(got, want):
    want = want.numpy()
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-6 * max(float(np.abs(want).max()), 1e-30))


@pytest.mark.parametrize("shape", [(2, 19, 32, 64), (1, 19, 33, 47), (3, 7, 16, 20), (1, 19, 128, 256), (2, 19, 5)])
def test_energy_backward(shape):
    from multishiftseg_b200 import deeplab
    g = torch.Generator().manual_seed(sum(shape))
    x = 3.0 * torch.randn(shape, generator=g)
    go = torch.randn((shape[0],) + shape[2:], generator=g)
    xc = x.clone().requires_grad_(True)
    so.energy_func(xc).backward(go)
    xg = x.cuda().requires_grad_(True)
    s = deeplab.energy_func(xg)
    assert s.requires_grad
    s.backward(go.cuda())
    close(xg.grad, xc.grad)


@pytest.mark.parametrize("align", [True, False])
@pytest.mark.parametrize("h,w,H,W", [(16, 32, 32, 64), (17, 23, 40, 51), (540 // 4, 960 // 4, 1080 // 4, 1920 // 4), (8, 8, 8, 8),
                                     (5, 7, 20, 28), (12, 10, 7, 5), (1, 9, 4, 30), (6, 1, 1, 1)])
def test_upsample_backward_is_the_adjoint(align, h, w, H, W):
    from multishiftseg_b200 import deeplab
    g = torch.Generator().manual_seed(h * 1000 + W)
    x = torch.randn((2, 3, h, w), generator=g)
    go = torch.randn((2, 3, H, W), generator=g)
    xc = x.clone().requires_grad_(True)
    torch.nn.functional.interpolate(xc, size=(H, W), mode="bilinear", align_corners=align).backward(go)
    xg = x.cuda().requires_grad_(True)
    deeplab.Upsample(xg, (H, W), align_corners=align).backward(go.cuda())
    close(xg.grad, xc.grad)
    # adjoint identity <A x, g> == <x, A^T g> with this library's own forward
    lhs = float((deeplab.Upsample(x.cuda(), (H, W), align_corners=align).double() * go.cuda().double()).sum())
    rhs = float((x.cuda().double() * xg.grad.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0)


@pytest.mark.parametrize("B,C,h,w,H,W", [(2, 19, 32, 64, 64, 128), (1, 19, 135, 240, 270, 480), (1, 5, 9, 11, 30, 17)])
def test_anomaly_score_backward_fused(B, C, h, w, H, W):
    """deepv3.py:283 end to end, and a downstream loss shaped like lib/loss.py's use of the score map (masked means)."""
    from multishiftseg_b200 import deeplab
    g = torch.Generator().manual_seed(B + C + h)
    x = 2.0 * torch.randn((B, C, h, w), generator=g)
    mask = torch.rand((B, H, W), generator=g) < 0.3

    def loss_of(score, m):
        return score[m].mean() - 0.5 * score[~m].mean() + (score ** 2).mean() * 0.01

    xc = x.clone().requires_grad_(True)
    loss_of(so.deeplab_anomaly_score(xc, (H, W)), mask).backward()
    xg = x.cuda().requires_grad_(True)
    loss_of(deeplab.anomaly_score(xg, (H, W)), mask.cuda()).backward()
    close(xg.grad, xc.grad)


def test_no_grad_paths_unchanged():
    from multishiftseg_b200 import deeplab
    x = torch.randn((1, 19, 8, 16)).cuda().requires_grad_(True)
    with torch.no_grad():
        assert not deeplab.energy_func(x).requires_grad
        assert not deeplab.anomaly_score(x, (16, 32)).requires_grad
    assert not deeplab.energy_func(x.detach()).requires_grad


# ---- Mask2Former half: backward of the fused anomaly score -------------------------------------------
def _m2f_ref_grads(cls, lo, padded, crop, g):
    """torch autograd (CPU, fp32) through the oracle's restatement of maskformer_model.py:271-277 + train_m2f.py:387-407"""
    from oracle import scoring_oracle as so
    c = cls.clone().requires_grad_(True)
    m = lo.clone().requires_grad_(True)
    a = so.m2f_anomaly_from_lowres(c, m, padded, crop)
    a.backward(g)
    return a.detach(), c.grad, m.grad


@pytest.mark.parametrize("B,Q,hw,crop", [(1, 100, (16, 32), (64, 128)), (2, 100, (16, 32), (61, 125)), (1, 37, (9, 14), (36, 56)),
                                         (1, 8, (12, 20), (40, 70))])
def test_m2f_anomaly_backward_matches_torch_autograd(B, Q, hw, crop):
    from multishiftseg_b200 import m2f
    g = torch.Generator().manual_seed(B * 1000 + Q)
    cls = 3.0 * torch.randn((B, Q, 20), generator=g)
    lo = 4.0 * torch.randn((B, Q) + hw, generator=g)
    padded = (4 * hw[0], 4 * hw[1])
    go = torch.randn((B,) + crop, generator=g)
    a_ref, gc_ref, gm_ref = _m2f_ref_grads(cls, lo, padded, crop, go)
    c = cls.cuda().requires_grad_(True)
    m = lo.cuda().requires_grad_(True)
    a = m2f.anomaly_score_from_lowres(c, m, padded, crop)
    np.testing.assert_allclose(a.detach().cpu().numpy(), a_ref.numpy(), rtol=1e-5, atol=2e-6)
    a.backward(go.cuda())
    for got, want in ((c.grad, gc_ref), (m.grad, gm_ref)):
        tol = 1e-5 * float(want.abs().max()) + 1e-7
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-4, atol=tol)


def test_m2f_anomaly_backward_non_x4_resize_and_get_anomaly_score():
    """generic resize factor (the adjoint is recomputed from the forward's own tap arithmetic) and the reference entry
    point get_anomaly_score with already-upsampled masks (identity resize)"""
    from multishiftseg_b200 import m2f
    from oracle import scoring_oracle as so
    g = torch.Generator().manual_seed(5)
    cls = 2.0 * torch.randn((1, 20, 20), generator=g)
    lo = 3.0 * torch.randn((1, 20, 10, 15), generator=g)
    go = torch.randn((1, 25, 38), generator=g)
    a_ref, gc_ref, gm_ref = _m2f_ref_grads(cls, lo, (25, 38), (25, 38), go)
    c, m = cls.cuda().requires_grad_(True), lo.cuda().requires_grad_(True)
    m2f.anomaly_score_from_lowres(c, m, (25, 38), (25, 38)).backward(go.cuda())
    np.testing.assert_allclose(c.grad.cpu().numpy(), gc_ref.numpy(), rtol=1e-4, atol=1e-5 * float(gc_ref.abs().max()))
    np.testing.assert_allclose(m.grad.cpu().numpy(), gm_ref.numpy(), rtol=1e-4, atol=1e-5 * float(gm_ref.abs().max()))
    # get_anomaly_score: masks already at full resolution
    up = 3.0 * torch.randn((1, 20, 24, 40), generator=g)
    go2 = torch.randn((1, 20, 33), generator=g)
    cr = cls.clone().requires_grad_(True)
    ur = up.clone().requires_grad_(True)
    so.get_anomaly_score({"pred_logits_ood": cr, "pred_masks_ood": ur}, (20, 33)).backward(go2)
    c2, u2 = cls.cuda().requires_grad_(True), up.cuda().requires_grad_(True)
    m2f.get_anomaly_score({"pred_logits_ood": c2, "pred_masks_ood": u2}, (20, 33)).backward(go2.cuda())
    np.testing.assert_allclose(c2.grad.cpu().numpy(), cr.grad.numpy(), rtol=1e-4, atol=1e-5 * float(cr.grad.abs().max()))
    np.testing.assert_allclose(u2.grad.cpu().numpy(), ur.grad.numpy(), rtol=1e-4, atol=1e-5 * float(ur.grad.abs().max()))


def test_m2f_anomaly_backward_is_deterministic():
    from multishiftseg_b200 import m2f
    g = torch.Generator().manual_seed(9)
    cls = (3.0 * torch.randn((2, 100, 20), generator=g)).cuda()
    lo = (4.0 * torch.randn((2, 100, 32, 64), generator=g)).cuda()
    go = torch.randn((2, 128, 256), generator=g).cuda()
    outs = []
    for _ in range(2):
        c, m = cls.clone().requires_grad_(True), lo.clone().requires_grad_(True)
        m2f.anomaly_score_from_lowres(c, m, (128, 256), (128, 256)).backward(go)
        outs.append((c.grad.clone(), m.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
