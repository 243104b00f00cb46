"""GPU parity of the confusion-histogram kernels (lib/utils/metric.py:10-64) against the
reference-generated fixtures and the numpy oracle.  Integer work: bit-exact."""
import json
import os
import warnings

import numpy as np
import pytest
import torch

import gen_inputs as gi
from multishiftseg_b200 import segmetric as sm
from oracle import segmetric_oracle as so

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "segmetric_golden.json")))


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"seed{c['seed']}-{c['mode']}-n{c['n']}")
@pytest.mark.parametrize("dtype", [torch.int64, torch.uint8])
def test_hist_info_golden(case, dtype):
    pred, gt = gi.confusion_case(case["seed"], case["n"], case["n_cl"], case["mode"])
    if dtype == torch.uint8:
        gt = np.where(gt < 0, 255, gt)                      # -1 and 255 are both "unlabeled"
    p = torch.from_numpy(pred).to(dtype).cuda()
    g = torch.from_numpy(gt).to(dtype).cuda()
    hist, labeled, correct = sm.hist_info(case["n_cl"], p, g)
    assert hist.dtype == np.int64 and hist.shape == (19, 19)
    assert hist.reshape(-1).tolist() == case["hist"]
    assert (int(labeled), int(correct)) == (case["labeled"], case["correct"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = sm.compute_metric([{"hist": hist, "labeled": labeled, "correct": correct}])
    assert float(a[0]).hex() == case["mean_IU"] and float(a[1]).hex() == case["mean_pixel_acc"]


def test_numpy_inputs_and_mixed_dtypes():
    pred, gt = gi.confusion_case(32, 1000, 19, "rand")
    want = so.hist_info(19, pred, gt)
    got = sm.hist_info(19, pred.astype(np.int32), gt)          # numpy in, int32 pred + int64 gt
    assert got[0].tolist() == want[0].tolist() and (int(got[1]), int(got[2])) == want[1:]


@pytest.mark.parametrize("case", GOLD["logits_cases"], ids=lambda c: f"seed{c['seed']}")
def test_fused_logits_golden(case):
    B, C, H, W = case["shape"]
    x, gt = gi.confusion_logits_case(case["seed"], B, C, H, W)
    acc = sm.ConfusionAccumulator(C, "cuda")
    acc.update_from_logits(torch.from_numpy(x).cuda(), torch.from_numpy(gt).cuda())
    hist, labeled, correct = acc.result()
    assert hist.reshape(-1).tolist() == case["hist"]
    assert (int(labeled), int(correct)) == (case["labeled"], case["correct"])


@pytest.mark.parametrize("gt_dtype", [torch.int64, torch.int32, torch.uint8])
def test_fused_logits_vs_oracle_with_ties_and_nan(gt_dtype):
    x, gt = gi.confusion_logits_case(7, 2, 19, 64, 128)
    x[0, 3, 5, 7] = np.nan                                      # torch.argmax: NaN is maximal
    x[1, :, 9, 9] = 0.0                                         # all classes tie -> class 0
    xt = torch.from_numpy(x)
    pred = xt.argmax(1).numpy()
    assert pred[0, 5, 7] == 3 and pred[1, 9, 9] == 0
    want = so.hist_info(19, pred, gt)
    acc = sm.ConfusionAccumulator(19, "cuda")
    acc.update_from_logits(xt.cuda(), torch.from_numpy(gt).to(gt_dtype).cuda())
    got = acc.result()
    assert got[0].tolist() == want[0].tolist() and (int(got[1]), int(got[2])) == want[1:]


def test_streaming_accumulation_equals_one_shot():
    """compute_metric's `hist += d['hist']` loop (metric.py:27-33) kept on the device across batches."""
    pred, gt = gi.confusion_case(36, 1 << 20, 19, "blobs")
    acc = sm.ConfusionAccumulator(19, "cuda")
    for part in range(4):
        s = slice(part << 18, (part + 1) << 18)
        acc.update(torch.from_numpy(pred[s]).cuda(), torch.from_numpy(gt[s]).cuda())
    hist, labeled, correct = acc.result()
    case = next(c for c in GOLD["cases"] if c["seed"] == 36)
    assert hist.reshape(-1).tolist() == case["hist"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = acc.compute()
    assert float(a[0]).hex() == case["mean_IU"] and float(a[1]).hex() == case["mean_pixel_acc"]


def test_full_size_properties():
    """cfg-2 sized batch (4 x 19 x 1024 x 2048): sum(hist) == labeled, trace == correct, == torch argmax + bincount."""
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((4, 19, 1024, 2048), device="cuda", generator=g)
    gt = torch.randint(0, 19, (4, 1024, 2048), device="cuda", generator=g)
    gt[torch.rand((4, 1024, 2048), device="cuda", generator=g) < 0.07] = 255
    acc = sm.ConfusionAccumulator(19, "cuda")
    acc.update_from_logits(x, gt)
    hist, labeled, correct = acc.result()
    assert hist.sum() == labeled == int((gt != 255).sum())
    assert np.trace(hist) == correct
    pred = x.argmax(1)
    k = gt != 255
    ref = torch.bincount(19 * gt[k] + pred[k], minlength=361).cpu().numpy().reshape(19, 19)
    assert hist.tolist() == ref.tolist()
    # class-index path on the same data
    h2, l2, c2 = sm.hist_info(19, pred, gt)
    assert h2.tolist() == ref.tolist() and (l2, c2) == (labeled, correct)


def test_out_of_range_pred_raises_like_numpy():
    with pytest.raises(ValueError):
        sm.hist_info(19, torch.tensor([25], device="cuda"), torch.tensor([18], device="cuda"))
    with pytest.raises(AssertionError):
        sm.hist_info(19, torch.zeros(3, device="cuda", dtype=torch.int64), torch.zeros(4, device="cuda", dtype=torch.int64))


def test_cpu_tensors_without_gpu_path_rejected_for_logits():
    from multishiftseg_b200 import _lib as L
    acc = sm.ConfusionAccumulator(19, "cuda")
    with pytest.raises(L.MssError):
        acc.update_from_logits(torch.zeros(1, 19, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64))
