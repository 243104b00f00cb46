"""The mIoU-half oracle against fixtures produced by executing the reference's own hist_info / compute_metric
(tests/golden/make_golden_segmetric.py).  CPU only."""
import json
import os
import warnings

import numpy as np
import pytest

import gen_inputs as gi
from oracle import segmetric_oracle as so

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "segmetric_golden.json")))


def check_floats(case, hist, labeled, correct, cm):
    res = [{"hist": hist, "labeled": labeled, "correct": correct}]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = cm(res)
        b = cm(res, per_class=True)
    assert float(a[0]).hex() == case["mean_IU"] and float(a[1]).hex() == case["mean_pixel_acc"]
    assert float(b[0]).hex() == case["pc_mean_IU"] and float(b[1]).hex() == case["pc_mean_pixel_acc"]
    assert [float(x).hex() for x in b[2]] == case["pc_iu"]
    assert [float(x).hex() for x in b[3]] == case["pc_class_acc"]


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"seed{c['seed']}-{c['mode']}-n{c['n']}")
def test_hist_info_matches_reference(case):
    pred, gt = gi.confusion_case(case["seed"], case["n"], case["n_cl"], case["mode"])
    assert gi.digest(pred, gt) == case["sha256"], "generated inputs drifted from the fixture"
    hist, labeled, correct = so.hist_info(case["n_cl"], pred, gt)
    assert hist.reshape(-1).tolist() == case["hist"]
    assert (labeled, correct) == (case["labeled"], case["correct"])
    check_floats(case, hist, labeled, correct, so.compute_metric)


@pytest.mark.parametrize("case", GOLD["logits_cases"], ids=lambda c: f"seed{c['seed']}")
def test_logits_argmax_matches_reference(case):
    B, C, H, W = case["shape"]
    x, gt = gi.confusion_logits_case(case["seed"], B, C, H, W)
    assert gi.digest(x, gt) == case["sha256"]
    hist, labeled, correct = so.hist_info(C, so.argmax_first(x), gt)
    assert hist.reshape(-1).tolist() == case["hist"]
    assert (labeled, correct) == (case["labeled"], case["correct"])


def test_out_of_range_pred_raises():
    with pytest.raises(ValueError):
        so.hist_info(19, np.array([25]), np.array([18]))


def test_host_mirror_floats_match_reference():
    """multishiftseg_b200.segmetric's closed-form 19 x 19 arithmetic (host numpy) == the reference's."""
    from multishiftseg_b200 import segmetric as sm
    for case in GOLD["cases"]:
        hist = np.asarray(case["hist"], dtype=np.int64).reshape(19, 19)
        check_floats(case, hist, case["labeled"], case["correct"], sm.compute_metric)
