"""Generate the golden fixtures by EXECUTING THE REFERENCE'S OWN CODE.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

* ``lib/utils/metric.py`` imports only numpy / sklearn / time, so it is loaded
  by file path and ``eval_ood_measure`` is run unmodified.
* ``deepv3.py`` / ``mynn.py`` / ``maskformer_model.py`` / ``train_m2f.py`` import
  packages that are not installed (easydict, detectron2, fvcore ...), so the
  hot-path *functions* are pulled out of those files with ``ast`` and compiled
  as they stand -- the reference source is read at generation time, never
  copied into this repo.

Outputs: ``tests/golden/metrics_golden.json`` (float64 as hex) and
``tests/golden/scoring_golden.npz``.  The GPU box has no /root/reference; tests
read only these fixtures.
"""
from __future__ import annotations

import ast
import importlib.util
import json
import os
import sys
import types
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_inputs as gi  # noqa: E402

REF = os.environ.get("MSS_REFERENCE", "/root/reference")


def load_by_path(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def extract_function(rel, func, cls=None):
    """Compile one def out of a reference file without importing the module."""
    path = os.path.join(REF, rel)
    tree = ast.parse(open(path).read(), filename=path)
    body = tree.body
    if cls is not None:
        body = next(n for n in body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == func)
    fn.decorator_list = []
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch, "nn": nn, "F": F, "Dict": Dict, "Tuple": Tuple}
    exec(compile(mod, path, "exec"), ns)
    return ns[func]


def hexes(t):
    return [float(x).hex() for x in t]


def main():
    import sklearn
    import scipy
    meta = {"numpy": np.__version__, "sklearn": sklearn.__version__, "scipy": scipy.__version__,
            "torch": torch.__version__, "reference": REF}

    # ------------------------------------------------------------------ metrics
    ref_metric = load_by_path("ref_metric", "lib/utils/metric.py")
    out = {"meta": meta, "kats": {}, "cases": []}
    for name, (s, l) in gi.KATS.items():
        s = np.asarray(s, dtype=np.float32)
        l = np.asarray(l, dtype=np.int64)
        try:
            r = ref_metric.eval_ood_measure(s, l)
            out["kats"][name] = None if r is None else hexes(r)
        except ValueError as e:
            out["kats"][name] = {"error": "ValueError", "message": str(e)}
    for (seed, n, mode, p_ood, p_ign) in gi.METRIC_CASES:
        s, l = gi.metric_case(seed, n, mode, p_ood, p_ign)
        r = ref_metric.eval_ood_measure(s, l)
        out["cases"].append({"seed": seed, "n": n, "mode": mode, "p_ood": p_ood, "p_ignore": p_ign,
                             "sha256": gi.digest(s, l),
                             "expected": None if r is None else hexes(r)})
        print("metric case", seed, n, mode, None if r is None else [float(x) for x in r])
    # a 2-D (image-shaped) call exactly as test_deeplab.py:98-101 makes it
    s, l = gi.metric_case(99, 2 * 64 * 128, "cont")
    r = ref_metric.eval_ood_measure(s.reshape(2, 64, 128), l.reshape(2, 64, 128))
    out["image_shaped"] = {"seed": 99, "shape": [2, 64, 128], "sha256": gi.digest(s, l), "expected": hexes(r)}
    with open(os.path.join(HERE, "metrics_golden.json"), "w") as f:
        json.dump(out, f, indent=1)

    # ------------------------------------------------------------------ scoring
    energy_func = extract_function("lib/network/deepv3/deepv3.py", "energy_func", cls="DeepWV3Plus")
    Upsample = extract_function("lib/network/deepv3/mynn.py", "Upsample")
    semantic_inference = extract_function("lib/network/mask2former/maskformer_model.py",
                                          "semantic_inference", cls="MaskFormer")
    get_anomaly_score = extract_function("train_m2f.py", "get_anomaly_score", cls="TrainM2FOOD")

    d = {k: torch.from_numpy(v) for k, v in gi.scoring_inputs().items()}
    g = {}
    with torch.no_grad():
        # a1: DeepWV3Plus.energy_func (deepv3.py:251-253)
        g["dl_energy"] = energy_func(None, d["dl_logit"])
        g["dl_energy_odd"] = energy_func(None, d["dl_logit_odd"])
        # a3: mynn.Upsample (mynn.py:28-33), x2 and a non-integer factor
        g["up_x2"] = Upsample(d["up_in"], (24, 40))
        g["up_odd"] = Upsample(d["up_in"], (31, 53))
        # deepv3.py:283: Upsample(energy_func(dec2).unsqueeze(1), size).squeeze(1)
        g["dl_anomaly_x2"] = Upsample(energy_func(None, d["dl_logit"]).unsqueeze(1), (48, 80)).squeeze(1)
        # a4: F.interpolate(..., mode="bilinear", align_corners=False) (maskformer_model.py:264-277)
        up = F.interpolate(d["m2f_mask_lo"], size=(32, 64), mode="bilinear", align_corners=False)
        g["m2f_mask_up_q8"] = up[:, :8].contiguous()   # first 8 queries only, keeps the fixture small
        # a5: MaskFormer.semantic_inference (maskformer_model.py:341-354)
        fake_self = types.SimpleNamespace(sem_seg_head=types.SimpleNamespace(num_classes=19))
        for b in range(2):
            g[f"m2f_semseg_{b}"] = semantic_inference(fake_self, d["m2f_cls"][b], up[b])
        # a7: TrainM2FOOD.get_anomaly_score (train_m2f.py:387-407), incl. the crop
        g["m2f_anomaly"] = get_anomaly_score(None, {"pred_logits_ood": d["m2f_cls"], "pred_masks_ood": up}, (30, 61))
    arrays = {k: v.numpy() for k, v in g.items()}
    arrays["inputs_sha256"] = np.frombuffer(
        gi.digest(*[v.numpy() for v in d.values()]).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "scoring_golden.npz"), **arrays)
    for k, v in arrays.items():
        print("scoring", k, v.shape)
    print("K channels:", [arrays[f"m2f_semseg_{b}"].shape[0] - 19 for b in range(2)])


if __name__ == "__main__":
    main()
