"""The oracle is pinned here: every restatement in oracle/ must reproduce, bit for
bit, what the reference's own lib/utils/metric.py produced (tests/golden/
metrics_golden.json, made by tests/golden/make_golden.py) and the SURVEY 8(c)
known-answer tests K1-K13."""
import json
import os

import numpy as np
import pytest

import gen_inputs as gi
from oracle import c_oracle, metrics_oracle as mo

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "metrics_golden.json")))

# SURVEY.md section 8(c), exact float64 as hex
KAT_EXPECT = {
    "K1": (0.75, float.fromhex("0x1.aaaaaaaaaaaaap-1"), 0.5),
    "K3": (0.5, float.fromhex("0x1.5555555555555p-2"), 1.0),
    "K4": (0.84375, float.fromhex("0x1.9555555555555p-1"), 0.5),
    "K5": (1.0, 1.0, 0.0),
    "K6": (0.0, float.fromhex("0x1.aaaaaaaaaaaaap-2"), 1.0),
    "K9": (float.fromhex("0x1.8e38e38e38e39p-1"), float.fromhex("0x1.7777777777778p-1"), 0.6666666666666666),
    "K12": (1.0, 1.0, 0.0),
    "K13": (float.fromhex("0x1.f51b3bea3677ep-1"), float.fromhex("0x1.82d82d82d82d8p-1"),
            float.fromhex("0x1.5c9882b931057p-5")),
}
KAT_EXPECT["K2"] = KAT_EXPECT["K1"]

IMPLS = {
    "sklearn_flow": mo.eval_ood_measure,
    "counts_numpy": mo.eval_ood_measure_counts,
    "counts_c": c_oracle.eval_ood_measure,
}


def _unhex(v):
    return None if v is None else tuple(float.fromhex(x) for x in v)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", sorted(gi.KATS))
def test_kats(impl, name):
    s, l = gi.KATS[name]
    s = np.asarray(s, dtype=np.float32)
    l = np.asarray(l, dtype=np.int64)
    gold = GOLD["kats"][name]
    if isinstance(gold, dict):
        with pytest.raises(ValueError) as ei:
            IMPLS[impl](s, l)
        assert str(ei.value) == gold["message"]
        return
    r = IMPLS[impl](s, l)
    if gold is None:
        assert r is None
        return
    r = tuple(float(x) for x in r)
    assert r == _unhex(gold)
    assert r == KAT_EXPECT[name]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"s{c['seed']}-n{c['n']}-{c['mode']}")
def test_golden_cases(impl, case):
    if impl == "counts_numpy" and case["n"] > 300000:
        pytest.skip("pure-numpy leaf loop is slow; the C restatement covers the large cases")
    s, l = gi.metric_case(case["seed"], case["n"], case["mode"], case["p_ood"], case["p_ignore"])
    assert gi.digest(s, l) == case["sha256"], "synthetic input drifted from what the reference was run on"
    r = IMPLS[impl](s, l)
    exp = _unhex(case["expected"])
    if exp is None:
        assert r is None
    else:
        assert tuple(float(x) for x in r) == exp


def test_image_shaped_call():
    c = GOLD["image_shaped"]
    s, l = gi.metric_case(c["seed"], int(np.prod(c["shape"])), "cont")
    assert gi.digest(s, l) == c["sha256"]
    for f in IMPLS.values():
        r = f(s.reshape(c["shape"]), l.reshape(c["shape"]))
        assert tuple(float(x) for x in r) == _unhex(c["expected"])


@pytest.mark.parametrize("ldt", ["uint8", "int32", "int64"])
def test_c_oracle_label_dtypes(ldt):
    s, l = gi.metric_case(7, 1000, "q2", label_dtype=ldt)
    exp = _unhex(next(c for c in GOLD["cases"] if c["seed"] == 7)["expected"])
    assert c_oracle.eval_ood_measure(s, l) == exp


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 127, 128, 129, 255, 256, 257, 1000, 8191, 65537, 1000003])
def test_pairwise_sum_is_numpy_sum(n):
    a = np.random.default_rng(n).standard_normal(n) * 1e3
    assert c_oracle.pairwise_sum(a) == float(np.sum(a))
    if n <= 70000:
        assert mo.pairwise_sum(a) == float(np.sum(a))
        leaves = mo.pairwise_leaves(n)
        assert sum(m for _, m in leaves) == n
        assert all(m <= 128 for _, m in leaves)


def test_counts_invariants():
    s, l = gi.metric_case(6, 1000, "q2")
    tps, fps = mo.ood_counts(s, l)
    assert tps[-1] == (l == 1).sum() and fps[-1] == (l == 0).sum()
    assert (np.diff(tps) >= 0).all() and (np.diff(fps) >= 0).all()
    assert (np.diff(tps + fps) > 0).all()
    assert c_oracle.metrics_from_counts(tps, fps) == mo.metrics_from_counts(tps, fps)


def test_key_roundtrip_and_order():
    x = np.array([-np.float32(3.5), -0.0, 0.0, 1e-45, 1.0, 2.5, -1e-45, 3.4e38, -3.4e38], dtype=np.float32)
    k = mo.float_key_desc(x)
    assert k[1] == k[2]
    back = mo.key_to_float(k)
    assert np.array_equal(np.abs(back), np.abs(x)) and np.array_equal(back[[0, 3, 4, 5]], x[[0, 3, 4, 5]])
    order = np.argsort(k, kind="stable")
    assert (np.diff(x[order]) <= 0).all()
