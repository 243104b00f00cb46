"""Pins oracle/scoring_oracle.py against outputs of the reference's own function
bodies (tests/golden/scoring_golden.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

import gen_inputs as gi
from oracle import scoring_oracle as so

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "scoring_golden.npz"))
D = {k: torch.from_numpy(v) for k, v in gi.scoring_inputs().items()}


def test_inputs_are_the_ones_the_reference_saw():
    want = bytes(G["inputs_sha256"]).decode()
    assert gi.digest(*[v.numpy() for v in D.values()]) == want


def eq(a, name):
    """Same torch build, same CPU kernels: the restatement is expected to be bit-identical to what the
    reference's function body returned.  torch's CPU transcendental kernels were seen (once in many runs,
    not reproducible) to return a few-ulp different logsumexp inside a long pytest process, so a mismatch
    is only accepted inside the path's own tolerance (north_star: fp32 scores within 1e-5 relative) and is
    reported."""
    got, want = a.numpy(), G[name]
    assert got.shape == want.shape and got.dtype == want.dtype, name
    if np.array_equal(got, want):
        return
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-5, err_msg=name)
    import warnings
    warnings.warn(f"{name}: restatement within 1e-5 of the reference output but not bit-identical on this host "
                  f"({int((got != want).sum())} of {got.size} elements differ)")


def test_energy_func():
    eq(so.energy_func(D["dl_logit"]), "dl_energy")
    eq(so.energy_func(D["dl_logit_odd"]), "dl_energy_odd")


def test_upsample_align_corners_true():
    eq(so.Upsample(D["up_in"], (24, 40)), "up_x2")
    eq(so.Upsample(D["up_in"], (31, 53)), "up_odd")
    eq(so.deeplab_anomaly_score(D["dl_logit"], (48, 80)), "dl_anomaly_x2")


def test_m2f_chain():
    up = so.upsample_masks(D["m2f_mask_lo"], (32, 64))
    eq(up[:, :8].contiguous(), "m2f_mask_up_q8")
    for b in range(2):
        eq(so.semantic_inference(D["m2f_cls"][b], up[b]), f"m2f_semseg_{b}")
    eq(so.get_anomaly_score({"pred_logits_ood": D["m2f_cls"], "pred_masks_ood": up}, (30, 61)), "m2f_anomaly")
    eq(so.m2f_anomaly_from_lowres(D["m2f_cls"], D["m2f_mask_lo"], (32, 64), (30, 61)), "m2f_anomaly")
    for b, r in enumerate(so.m2f_post_head(D["m2f_cls"], D["m2f_mask_lo"], (32, 64), (32, 64))):
        eq(r, f"m2f_semseg_{b}")


def test_extra_scores_consistent():
    """a2 has no reference code ("parity unpinned"): check the definitions agree with each other."""
    x = D["dl_logit"]
    p = torch.softmax(x.double(), dim=1)
    assert torch.allclose(so.msp_score(x).double(), 1 - p.max(1)[0], atol=1e-6)
    assert torch.allclose(so.entropy_score(x).double(), -(p * p.log()).sum(1), atol=1e-5)
    assert torch.equal(so.maxlogit_score(x), -x.max(1)[0])
    # energy <= maxlogit score (logsumexp >= max)
    assert (so.energy_func(x) <= so.maxlogit_score(x) + 1e-6).all()
