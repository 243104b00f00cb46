"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports exactly what
include/mss_b200.h declares; the product package contains no route to the oracle or to a CPU fallback."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mss_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"^MSS_API[^;(]*?\b(mss_[a-z0-9_]+)\s*\(", src, flags=re.M)))


@pytest.fixture(scope="module")
def lib_path():
    from multishiftseg_b200 import build
    return build.build()


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ["mss_deeplab_score", "mss_upsample_bilinear", "mss_m2f_semantic_inference", "mss_ood_metrics",
                 "mss_sort_keys", "mss_eval_sort", "mss_counts_from_sorted", "mss_partition_scatter_keys", "mss_metrics_tail", "mss_eval_append", "mss_last_error"]:
        assert must in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in mss_b200.h but not exported"


def test_ctypes_table_matches_header(lib_path):
    from multishiftseg_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.mss_abi_version() == 2
    # pure host-side size queries are callable without a GPU
    assert lib.mss_sort_keys_workspace_bytes(1 << 20) > (1 << 20) * 4
    assert lib.mss_ood_metrics_workspace_bytes(1000) > 0
    assert lib.mss_tail_workspace_bytes(10) > 0
    assert lib.mss_m2f_workspace_bytes(2, 100, 19) >= 2 * 100 * 20 * 4
    assert lib.mss_deeplab_score_host_scratch_bytes(16, 19, 1024 * 2048, 11) > 3 * 19 * 1024 * 2048 * 4


def test_only_sm100a_code_in_library(lib_path):
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_never_touches_oracle_or_cpu_fallback():
    pkg = os.path.join(ROOT, "multishiftseg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for banned in ("oracle", "sklearn", "scipy", "triton"):
                    assert not re.search(rf"^\s*(from|import)\s+{banned}\b", src, flags=re.M), (f, banned)
                assert "torch.compile(" not in src, f


def test_missing_cuda_raises_instead_of_falling_back():
    import numpy as np
    import torch
    from multishiftseg_b200 import _lib, deeplab, metric
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MssError):
        metric.eval_ood_measure(np.zeros(4, np.float32), np.array([0, 1, 0, 1]))
    with pytest.raises(_lib.MssError):
        deeplab.energy_func(torch.zeros(1, 19, 4, 4))
    # every entry point of the Python mirror refuses CPU tensors (no silent PyTorch fallback), with or without autograd
    from multishiftseg_b200 import m2f, segmetric
    x = torch.zeros(1, 19, 4, 4)
    calls = [
        lambda: deeplab.energy_func(x.clone().requires_grad_(True)),
        lambda: deeplab.anomaly_score(x, (8, 8)),
        lambda: deeplab.anomaly_score(x.clone().requires_grad_(True), (8, 8)),
        lambda: deeplab.Upsample(x, (8, 8)),
        lambda: deeplab.score_maps(x, ("energy", "entropy")),
        lambda: deeplab.head_scores(torch.zeros(1, 32, 4, 4), torch.zeros(19, 32), torch.zeros(19, 32)),
        lambda: m2f.mask_logits(torch.zeros(1, 100, 32), torch.zeros(1, 32, 4, 4)),
        lambda: m2f.anomaly_score_from_lowres(torch.zeros(1, 100, 20), torch.zeros(1, 100, 4, 4), (16, 16), (16, 16)),
        lambda: m2f.anomaly_score_from_features(torch.zeros(1, 100, 20), torch.zeros(1, 100, 32), torch.zeros(1, 32, 4, 4),
                                                (16, 16), (16, 16)),
        lambda: m2f.semantic_inference(torch.zeros(100, 20), torch.zeros(100, 8, 8)),
        lambda: segmetric.ConfusionAccumulator(19),
    ]
    for i, call in enumerate(calls):
        with pytest.raises(_lib.MssError):
            call()


@pytest.mark.parametrize("n", [1, 7, 8, 128, 129, 1000, 4097, 100_003, 1_000_001, 12_345_678])
def test_pairwise_tree_descent_matches_numpy_tree(lib_path, n):
    """Host-only entry point: the table descent the device kernels use to find leaf i of numpy's
    pairwise tree must give the leaves the oracle enumerates (oracle/metrics_oracle.py:pairwise_leaves,
    itself pinned to np.sum by tests/test_oracle_metrics.py)."""
    import ctypes as C
    import random
    from multishiftseg_b200 import _lib as L
    from oracle import metrics_oracle as mo
    lib = L.load()
    want = mo.pairwise_leaves(n)
    s, m, nl = C.c_int64(), C.c_int64(), C.c_int64()
    rng = random.Random(n)
    picks = range(len(want)) if len(want) <= 2000 else sorted(
        {0, 1, len(want) - 2, len(want) - 1, *(rng.randrange(len(want)) for _ in range(2000))})
    for i in picks:
        assert lib.mss_pairwise_leaf_bounds(n, i, C.byref(s), C.byref(m), C.byref(nl)) == 0, L.last_error()
        assert nl.value == len(want)
        assert (s.value, m.value) == tuple(want[i])
    assert lib.mss_pairwise_leaf_bounds(n, len(want), C.byref(s), C.byref(m), C.byref(nl)) < 0


@pytest.mark.parametrize("n", [1, 5, 8, 127, 128, 129, 1000, 32768, 32769, 65_537, 1_000_003, 40_000_001])
def test_pairwise_plan_and_combine_equal_numpy_sum(lib_path, n):
    """plan (size table + frontier) -> leaf descent -> subtree combine -> top combine, run on the host through
    the code shared with the device kernels, must reproduce np.sum bit for bit (ill-conditioned input)."""
    import ctypes as C
    import numpy as np
    from multishiftseg_b200 import _lib as L
    rng = np.random.default_rng(n)
    a = (rng.standard_normal(n) * np.exp(rng.uniform(-20, 20, n))).astype(np.float64)
    out = C.c_double()
    assert L.load().mss_pairwise_sum_host(a.ctypes.data, n, C.byref(out)) == 0, L.last_error()
    assert out.value == float(np.sum(a))


def test_forward_only_entry_points_refuse_autograd_inputs():
    """m2f.* / head_scores / score_maps have no backward; with grad-tracked inputs they must raise instead of
    silently returning tensors without grad_fn (host-side check, runs before any device work)."""
    import torch
    from multishiftseg_b200 import _lib
    x = torch.zeros(3, requires_grad=True)
    with pytest.raises(_lib.MssError):
        _lib.forbid_grad("op", x)
    with torch.no_grad():
        _lib.forbid_grad("op", x)                      # inference: fine
    _lib.forbid_grad("op", x.detach(), None, 3)
