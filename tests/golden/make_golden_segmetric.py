"""Golden fixtures for the mIoU half of the reference's lib/utils/metric.py (:10-64), made by EXECUTING
THE REFERENCE'S OWN CODE (the module imports only numpy / sklearn / time, so it loads by file path).

    python tests/golden/make_golden_segmetric.py        # in the build container (/root/reference exists)

Output: tests/golden/segmetric_golden.json -- per case the 19 x 19 histogram, labeled, correct and the float64
results (hex) of compute_metric(per_class=False / True).  The GPU box has no /root/reference; tests read
only the fixture.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_inputs as gi  # noqa: E402
from make_golden import REF, hexes, load_by_path  # noqa: E402


def main():
    ref = load_by_path("ref_metric", "lib/utils/metric.py")
    out = {"meta": {"numpy": np.__version__, "reference": REF}, "cases": [], "logits_cases": []}
    warnings.simplefilter("ignore")                      # 0/0 for absent classes: nan, as in the reference

    def record(pred, gt, n_cl):
        hist, labeled, correct = ref.hist_info(n_cl, pred, gt)
        res = [{"hist": hist, "labeled": labeled, "correct": correct}]
        a = ref.compute_metric(res)
        b = ref.compute_metric(res, per_class=True)
        # the same data fed as two batches (compute_metric's accumulation loop)
        h = pred.size // 2
        parts = [dict(zip(("hist", "labeled", "correct"), ref.hist_info(n_cl, p, g)))
                 for p, g in ((pred.reshape(-1)[:h], gt.reshape(-1)[:h]), (pred.reshape(-1)[h:], gt.reshape(-1)[h:]))]
        a2 = ref.compute_metric(parts)
        assert [float(x).hex() for x in a2] == [float(x).hex() for x in a]
        return {"hist": hist.reshape(-1).tolist(), "labeled": int(labeled), "correct": int(correct),
                "mean_IU": float(a[0]).hex(), "mean_pixel_acc": float(a[1]).hex(),
                "pc_mean_IU": float(b[0]).hex(), "pc_mean_pixel_acc": float(b[1]).hex(),
                "pc_iu": hexes(b[2]), "pc_class_acc": hexes(b[3])}

    for (seed, n, n_cl, mode) in gi.CONFUSION_CASES:
        pred, gt = gi.confusion_case(seed, n, n_cl, mode)
        r = record(pred, gt, n_cl)
        r.update({"seed": seed, "n": n, "n_cl": n_cl, "mode": mode, "sha256": gi.digest(pred, gt)})
        out["cases"].append(r)
        print("confusion", seed, n, mode, float.fromhex(r["mean_IU"]), float.fromhex(r["mean_pixel_acc"]))
    import torch
    for (seed, B, C, H, W) in [(41, 2, 19, 24, 40), (42, 1, 19, 13, 27)]:
        x, gt = gi.confusion_logits_case(seed, B, C, H, W)
        pred = torch.from_numpy(x).argmax(1).numpy()     # what a caller of hist_info would pass for a logit map
        r = record(pred, gt, C)
        r.update({"seed": seed, "shape": [B, C, H, W], "sha256": gi.digest(x, gt)})
        out["logits_cases"].append(r)
    with open(os.path.join(HERE, "segmetric_golden.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
