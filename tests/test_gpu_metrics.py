"""GPU parity, metric stage: the CUDA path (through the C ABI) must be BIT-IDENTICAL to the reference's
result -- golden vectors made by running the reference's metric.py, KATs K1-K13, and the C oracle on
seeded random inputs.  Integer stages (sort, partition, histogram, counts) are compared exactly too."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

import gen_inputs as gi
from oracle import c_oracle, metrics_oracle as mo

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "metrics_golden.json")))


@pytest.fixture(scope="module")
def M():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no fallback)"
    from multishiftseg_b200 import metric
    return metric


def _unhex(v):
    return None if v is None else tuple(float.fromhex(x) for x in v)


def _tup(r):
    return None if r is None else tuple(float(x) for x in r)


@pytest.mark.parametrize("name", sorted(gi.KATS))
def test_kats(M, name):
    s, l = gi.KATS[name]
    s = np.asarray(s, dtype=np.float32)
    l = np.asarray(l, dtype=np.int64)
    gold = GOLD["kats"][name]
    if isinstance(gold, dict):
        with pytest.raises(ValueError) as ei:
            M.eval_ood_measure(s, l)
        assert str(ei.value) == gold["message"]
    else:
        assert _tup(M.eval_ood_measure(s, l)) == _unhex(gold)


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"s{c['seed']}-n{c['n']}-{c['mode']}")
def test_golden_cases(M, case):
    s, l = gi.metric_case(case["seed"], case["n"], case["mode"], case["p_ood"], case["p_ignore"])
    assert gi.digest(s, l) == case["sha256"]
    assert _tup(M.eval_ood_measure(s, l)) == _unhex(case["expected"])


def test_image_shaped_cuda_tensors(M):
    c = GOLD["image_shaped"]
    s, l = gi.metric_case(c["seed"], int(np.prod(c["shape"])), "cont")
    st = torch.from_numpy(s).reshape(c["shape"]).cuda()
    lt = torch.from_numpy(l).reshape(c["shape"]).cuda()
    assert _tup(M.eval_ood_measure(st, lt)) == _unhex(c["expected"])


@pytest.mark.parametrize("ldt", ["uint8", "int32", "int64", "int16"])
def test_label_dtypes(M, ldt):
    s, l = gi.metric_case(7, 1000, "q2", label_dtype=ldt)
    exp = _unhex(next(c for c in GOLD["cases"] if c["seed"] == 7)["expected"])
    assert _tup(M.eval_ood_measure(s, l)) == exp


@pytest.mark.parametrize("n,mode", [(1, "cont"), (2, "cont"), (3, "q0"), (4095, "cont"), (4096, "q2"), (4097, "cont"),
                                    (8191, "f16"), (65536, "cont"), (1_000_003, "cont"), (1_000_003, "q2"),
                                    (5_000_011, "cont"), (5_000_011, "f16"), (16_777_216, "cont")])
def test_random_vs_c_oracle(M, n, mode):
    s, l = gi.metric_case(1000 + n % 977, n, mode, label_dtype="uint8")
    exp = c_oracle.eval_ood_measure(s, l)
    got = _tup(M.eval_ood_measure(s, l))
    assert got == exp


def test_other_ids_and_all_ignored(M):
    s, l = gi.metric_case(5, 5000, "cont")
    l2 = np.where(l == 0, 7, np.where(l == 1, 3, l))
    assert _tup(M.eval_ood_measure(s, l2, train_id_in=7, train_id_out=3)) == c_oracle.eval_ood_measure(s, l)
    assert M.eval_ood_measure(s, np.full_like(l, 255)) is None
    assert M.eval_ood_measure(np.zeros(0, np.float32), np.zeros(0, np.int64)) is None


def test_nan_in_ignored_pixels_is_fine(M):
    s, l = gi.metric_case(6, 4000, "cont")
    exp = c_oracle.eval_ood_measure(s, l)
    s = s.copy()
    s[l == 255] = np.nan
    assert _tup(M.eval_ood_measure(s, l)) == exp


def test_float64_scores_rejected(M):
    with pytest.raises(TypeError):
        M.eval_ood_measure(np.zeros(4, np.float64), np.array([0, 1, 0, 1]))


def test_get_measures_and_fpr(M):
    s, l = gi.metric_case(8, 20000, "q2")
    exp = c_oracle.eval_ood_measure(s, l)
    assert _tup(M.get_measures(s[l == 1], s[l == 0])) == exp
    assert _tup(M.get_and_print_results(s[l == 1], s[l == 0])) == exp
    v = l != 255
    assert float(M.fpr_and_fdr_at_recall(l[v], s[v])) == exp[2]
    assert float(M.fpr_and_fdr_at_recall(np.where(l[v] == 1, 1, -1), s[v])) == exp[2]
    with pytest.raises(ValueError):
        M.fpr_and_fdr_at_recall(l, s)          # {0,1,255}: not binary


# ------------------------------------------------------------------------------------ integer stages
def _keys_t(k):
    return torch.from_numpy(np.ascontiguousarray(k).view(np.int32)).cuda()


def _keys_np(t):
    return t.cpu().numpy().view(np.uint32)


def test_sort_keys_exact(M):
    rng = np.random.default_rng(0)
    for n in [1, 2, 31, 4096, 4097, 100_000, 3_000_001]:
        k = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
        if n > 1000:
            k[: n // 3] &= np.uint32(0xFF)              # heavy ties / skewed high digits
        kt = _keys_t(k)
        M.sort_keys(kt, n)
        assert np.array_equal(_keys_np(kt), np.sort(k))


@pytest.mark.parametrize("na,nb", [(0, 5), (5, 0), (1, 1), (4096, 4096), (4097, 1), (100_003, 7_001), (2_000_001, 300_007),
                                   (7, 1_000_003)])
def test_sort_two_segments_one_launch_sequence(M, na, nb):
    """Both streams of an evaluation are sorted by one sequence of launches (tiles numbered over both segments,
    look-back confined to the segment); unaligned second segment."""
    rng = np.random.default_rng(na * 31 + nb)
    a = rng.integers(0, 2 ** 32, size=na, dtype=np.uint64).astype(np.uint32)
    b = (rng.integers(0, 2 ** 32, size=nb, dtype=np.uint64).astype(np.uint32) >> np.uint32(9)) | np.uint32(0x3F000000)
    buf = torch.full((na + nb + 7,), -1, dtype=torch.int32, device="cuda")
    ta, tb = buf[:na], buf[na + 3: na + 3 + nb]         # the second array starts 4-byte aligned only
    ta.copy_(_keys_t(a)); tb.copy_(_keys_t(b))
    M.sort_keys(ta, na, tb, nb)
    assert np.array_equal(_keys_np(ta), np.sort(a)) and np.array_equal(_keys_np(tb), np.sort(b))
    assert (buf[na: na + 3] == -1).all() and (buf[na + 3 + nb:] == -1).all()      # nothing outside the arrays


@pytest.mark.parametrize("mode", ["low_byte_constant", "fp16_born", "all_equal", "top_byte_constant"])
def test_sort_single_bin_pass_skip(M, mode):
    """A digit that is the same for every key of a segment is skipped on the device (odd number of live passes:
    the result is copied back); compared with the un-skipped sort in a subprocess (MSS_SORT_NOSKIP=1)."""
    rng = np.random.default_rng(5)
    n = 1_000_003
    k = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    if mode == "low_byte_constant":
        k = (k & np.uint32(0xFFFFFF00)) | np.uint32(0x5A)
    elif mode == "fp16_born":
        s = rng.standard_normal(n).astype(np.float16).astype(np.float32)
        k = mo.float_key_desc(s)
    elif mode == "all_equal":
        k[:] = 0xDEADBEEF
    else:
        k = (k & np.uint32(0x00FFFFFF)) | np.uint32(0x42000000)
    kt = _keys_t(k)
    other = _keys_t(k[::-1].copy()[: n // 3])           # second segment with its own skip decisions
    M.sort_keys(kt, n, other, n // 3)
    assert np.array_equal(_keys_np(kt), np.sort(k))
    assert np.array_equal(_keys_np(other), np.sort(k[::-1][: n // 3]))


def test_sort_histogram_counter_fold():
    """The upfront digit histogram keeps 8-bit lane-private counters and folds them every HIST_EPOCH (<= 15) chunks;
    MSS_HIST_EPOCH=2 (read once per process) makes a 5 M-key input cross many more folds.  Unaligned key pointer too.
    MSS_SORT_NOSKIP=1: the same process also checks the sort with the pass skip disabled."""
    import subprocess
    import sys
    code = r"""
import numpy as np, torch, sys
sys.path.insert(0, ".")
from multishiftseg_b200 import metric as M
rng = np.random.default_rng(3)
n = 5_000_003
k = rng.integers(0, 2 ** 32, size=n + 1, dtype=np.uint64).astype(np.uint32)
k[: n // 2] &= np.uint32(0x0000FFFF)
k[n // 2: n // 2 + n // 4] |= np.uint32(0xFFFF0000)
kt = torch.from_numpy(k.view(np.int32)).cuda()[1:]          # 4-byte aligned, not 16
M.sort_keys(kt, n)
assert np.array_equal(kt.cpu().numpy().view(np.uint32), np.sort(k[1:]))
k2 = (k[1:] & np.uint32(0xFFFFFF00)).copy()                  # low byte constant, skip disabled by the env
kt2 = torch.from_numpy(k2.view(np.int32)).cuda()
M.sort_keys(kt2, n)
assert np.array_equal(kt2.cpu().numpy().view(np.uint32), np.sort(k2))
print("fold ok")
"""
    env = dict(os.environ, MSS_HIST_EPOCH="2", MSS_SORT_NOSKIP="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "fold ok" in r.stdout, r.stdout + r.stderr


def _streams(s, l):
    """oracle-side streams: sorted negative / positive keys of the valid pixels"""
    return np.sort(mo.float_key_desc(s[l == 0])), np.sort(mo.float_key_desc(s[l == 1]))


@pytest.mark.parametrize("n,mode,p_ood", [(300_000, "f16", 0.05), (1_000_003, "cont", 0.5), (50_000, "const", 0.3),
                                          (2049, "q2", 0.1), (2048, "cont", 0.0), (2047, "cont", 1.0), (5_000_011, "q2", 0.02)])
def test_counts_and_tail_stage_level(M, n, mode, p_ood):
    """merge-path counts over the two sorted streams == the integer spec (oracle.ood_counts), with and without the
    multi-GPU prefixes; the tail over them == the C oracle."""
    s, l = gi.metric_case(11, n, mode, min(max(p_ood, 0.1), 0.9), 0.05, label_dtype="uint8")
    if p_ood == 0.0:
        l[l == 1] = 0
    if p_ood == 1.0:
        l[l == 0] = 1
    neg, pos = _streams(s, l)
    nt, pt = _keys_t(neg), _keys_t(pos)
    tps, fps = M.counts_from_sorted(nt, neg.size, pt, pos.size)
    keys = np.unique(np.concatenate([neg, pos]))
    tps_ref = np.searchsorted(pos, keys, side="right")
    fps_ref = np.searchsorted(neg, keys, side="right")
    assert np.array_equal(tps.cpu().numpy(), tps_ref) and np.array_equal(fps.cpu().numpy(), fps_ref)
    if 0.0 < p_ood < 1.0:
        t2, f2 = mo.ood_counts(s, l)
        assert np.array_equal(tps_ref, t2) and np.array_equal(fps_ref, f2)
        # offsets (the multi-GPU path): global prefixes are added exactly
        tps2, fps2 = M.counts_from_sorted(nt, neg.size, pt, pos.size, pos_before=10, neg_before=90)
        assert np.array_equal(tps2.cpu().numpy(), tps_ref + 10) and np.array_equal(fps2.cpu().numpy(), fps_ref + 90)
        res, t_roc = M.metrics_tail(tps, fps)
        assert _tup(res) == c_oracle.metrics_from_counts(tps_ref, fps_ref)


def test_histogram_and_partition(M):
    from multishiftseg_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(3)
    n = 1_234_567
    k = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    kt = _keys_t(k)
    st = torch.cuda.current_stream().cuda_stream
    hist = torch.empty(1 << 16, dtype=torch.int64, device="cuda")
    assert lib.mss_keys_histogram(kt.data_ptr(), n, 16, hist.data_ptr(), st) == 0
    assert np.array_equal(hist.cpu().numpy(), np.bincount(k >> 16, minlength=1 << 16))
    spl = np.array([1 << 30, 1 << 31, 3 << 30], dtype=np.uint32)
    splt = _keys_t(spl)
    ko = torch.empty_like(kt)
    nb = lib.mss_partition_workspace_bytes(n, 4)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    counts = (C.c_int64 * 4)()
    assert lib.mss_partition_keys(kt.data_ptr(), n, splt.data_ptr(), 4, ko.data_ptr(), counts, ws.data_ptr(), nb, st) == 0, L.last_error()
    dest = np.searchsorted(spl, k, side="right")
    order = np.argsort(dest, kind="stable")
    assert list(counts) == np.bincount(dest, minlength=4).tolist()
    assert np.array_equal(_keys_np(ko), k[order])


@pytest.mark.parametrize("bits", [16, 15, 12, 3])
def test_keys_histogram_skewed_and_unaligned(M, bits):
    """Score-like keys: almost everything in a few hundred bins (the case that made global atomics crawl),
    runs of equal bins, an unaligned key pointer, sizes around the vector / chunk boundaries."""
    from multishiftseg_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(bits)
    st = torch.cuda.current_stream().cuda_stream
    for n in [1, 3, 5, 4099, 700_001, 3_000_003]:
        s = (rng.standard_normal(n + 1) * 2 - 3).astype(np.float32)
        s[: n // 2] = np.repeat(s[: (n // 2 + 63) // 64], 64)[: n // 2]        # runs of 64 equal scores
        u = s.view(np.uint32)
        k = ~np.where(u >> 31, ~u, u | np.uint32(0x80000000))
        kt = torch.from_numpy(k.view(np.int32)).cuda()[1:]                    # 4-byte aligned only
        hist = torch.empty(1 << bits, dtype=torch.int64, device="cuda")
        assert lib.mss_keys_histogram(kt.data_ptr(), n, bits, hist.data_ptr(), st) == 0
        assert np.array_equal(hist.cpu().numpy(), np.bincount(k[1:] >> (32 - bits), minlength=1 << bits))


def test_keys_histogram_sampled(M):
    from multishiftseg_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(9)
    n = 1_000_003
    k = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    kt = torch.from_numpy(k.view(np.int32)).cuda()                          # 16-byte aligned: groups start at key 0
    st = torch.cuda.current_stream().cuda_stream
    for every in [1, 2, 7, 64]:
        hist = torch.empty(1 << 16, dtype=torch.int64, device="cuda")
        assert lib.mss_keys_histogram_sampled(kt.data_ptr(), n, 16, every, hist.data_ptr(), st) == 0
        sample = k if every == 1 else k[: n - n % 4].reshape(-1, 4)[::every].reshape(-1)
        assert np.array_equal(hist.cpu().numpy(), np.bincount(sample >> 16, minlength=1 << 16))


def test_keys_histogram_refine(M):
    """Second level of the splitter histogram: low 16 bits of the keys under one top-16-bit prefix, same sampling."""
    from multishiftseg_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(10)
    n = 1_000_003
    k = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    k[: 700_000] = (k[: 700_000] & np.uint32(0xFFFF)) | np.uint32(0xBF7F0000)  # 70 % of the keys under one prefix
    k[: 100_000] = np.uint32(0xBF7F1234)                                       # and a tie mass inside it
    k = rng.permutation(k)
    kt = torch.from_numpy(k.view(np.int32)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for prefix in [0xBF7F, 0x0000, 0xFFFF, 0x1234]:
        for every in [1, 3, 64]:
            hist = torch.full((1 << 16,), -1, dtype=torch.int64, device="cuda")
            assert lib.mss_keys_histogram_refine(kt.data_ptr(), n, prefix, every, hist.data_ptr(), st) == 0
            sample = k if every == 1 else k[: n - n % 4].reshape(-1, 4)[::every].reshape(-1)
            sample = sample[(sample >> 16) == prefix]
            assert np.array_equal(hist.cpu().numpy(), np.bincount(sample & 0xFFFF, minlength=1 << 16)), (prefix, every)
    hist = torch.full((1 << 16,), -1, dtype=torch.int64, device="cuda")
    assert lib.mss_keys_histogram_refine(0, 0, 5, 1, hist.data_ptr(), st) == 0      # empty input: zeros
    assert int(hist.abs().sum()) == 0
    assert lib.mss_keys_histogram_refine(kt.data_ptr(), n, 1 << 16, 1, hist.data_ptr(), st) != 0
    assert lib.mss_keys_histogram_refine(kt.data_ptr(), n, 1, 0, hist.data_ptr(), st) != 0


@pytest.mark.parametrize("parts", [1, 2, 3, 8, 16, 17, 200, 256])
def test_partition_many_and_few_parts(M, parts):
    from multishiftseg_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(parts)
    n = 777_777
    k = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    k[: n // 2] = (k[: n // 2] >> 12) | np.uint32(0x40000000)               # half the keys in a narrow range
    k[:5] = 0xFFFFFFFF                                                      # the largest key (beyond every splitter)
    spl = np.sort(rng.choice(k, size=parts - 1, replace=False)).astype(np.uint32) if parts > 1 else np.zeros(0, np.uint32)
    kt = _keys_t(k)
    splt = _keys_t(np.concatenate([spl, np.zeros(1, np.uint32)]))
    ko = torch.empty_like(kt)
    nb = lib.mss_partition_workspace_bytes(n, parts)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    counts = (C.c_int64 * parts)()
    st = torch.cuda.current_stream().cuda_stream
    assert lib.mss_partition_keys(kt.data_ptr(), n, splt.data_ptr(), parts, ko.data_ptr(), counts, ws.data_ptr(), nb, st) == 0, L.last_error()
    dest = np.searchsorted(spl, k, side="right")
    order = np.argsort(dest, kind="stable")
    assert list(counts) == np.bincount(dest, minlength=parts).tolist()
    assert np.array_equal(_keys_np(ko), k[order])


@pytest.mark.parametrize("parts", [2, 8])
def test_partition_scatter_to_separate_buffers(M, parts):
    """mss_partition_count + mss_partition_scatter_keys with one buffer per bucket (on a multi-GPU box these
    are the peers' receive buffers; here they are local): == the stable partition, bucket by bucket."""
    from multishiftseg_b200.evaluator import CudaBackend
    be = CudaBackend("cuda")
    rng = np.random.default_rng(parts)
    n = 1_000_001
    k = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    spl = [int(x) for x in np.sort(rng.choice(k, size=parts - 1, replace=False))]
    kt = _keys_t(k)
    counts = be.partition_count(kt, n, spl, parts)
    dest = np.searchsorted(np.asarray(spl, dtype=np.uint32), k, side="right")
    assert counts == np.bincount(dest, minlength=parts).tolist()
    off = 5                                                           # this "rank's" block starts at element 5
    bk = [torch.full((c + off + 3,), -1, dtype=torch.int32, device="cuda") for c in counts]
    be.partition_scatter(kt, n, spl, parts, [t.data_ptr() for t in bk], [off] * parts)
    for d in range(parts):
        sel = dest == d
        assert np.array_equal(_keys_np(bk[d])[off:off + counts[d]], k[sel])      # stable
        assert (bk[d][:off] == -1).all() and (bk[d][off + counts[d]:] == -1).all()                    # nothing outside the block


def test_two_stream_buffer_layout_and_grow(M):
    """in-distribution keys grow upwards from 0, OOD keys downwards from the capacity; grow() keeps both."""
    s, l = gi.metric_case(21, 100_000, "cont", 0.1, 0.1, label_dtype="uint8")
    buf = M.PairBuffer(s.size, "cuda")
    buf.append(torch.from_numpy(s).cuda(), torch.from_numpy(l).cuda())
    m, n_pos, nan, inf = buf.read_state()
    assert (m, n_pos, nan, inf) == (int((l != 255).sum()), int((l == 1).sum()), 0, 0)
    neg, pos = buf.streams(m, n_pos)
    assert np.array_equal(np.sort(_keys_np(neg)), np.sort(mo.float_key_desc(s[l == 0])))
    assert np.array_equal(np.sort(_keys_np(pos)), np.sort(mo.float_key_desc(s[l == 1])))
    buf.grow(3 * s.size)
    buf.append(torch.from_numpy(s).cuda(), torch.from_numpy(l).cuda())
    assert _tup(M._finish(buf)) == c_oracle.eval_ood_measure(np.concatenate([s, s]), np.concatenate([l, l]))


def test_one_shot_c_entry(M):
    from multishiftseg_b200 import _lib as L
    lib = L.load()
    s, l = gi.metric_case(12, 777_777, "cont", label_dtype="int64")
    st_, lt = torch.from_numpy(s).cuda(), torch.from_numpy(l).cuda()
    nb = lib.mss_ood_metrics_workspace_bytes(s.size)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    out, cnt = (C.c_double * 3)(), (C.c_int64 * 4)()
    rc = lib.mss_ood_metrics(st_.data_ptr(), lt.data_ptr(), L.LABEL_I64, s.size, 0, 1, ws.data_ptr(), nb, out, cnt,
                             torch.cuda.current_stream().cuda_stream)
    assert rc == 0, L.last_error()
    exp, counts = c_oracle.eval_ood_measure(s, l, return_counts=True)
    assert tuple(out) == exp
    assert list(cnt) == counts.tolist()
    # empty class -> MSS_EMPTY_CLASS (reference: None)
    lt0 = torch.zeros_like(lt)
    assert lib.mss_ood_metrics(st_.data_ptr(), lt0.data_ptr(), L.LABEL_I64, s.size, 0, 1, ws.data_ptr(), nb, out, cnt,
                               torch.cuda.current_stream().cuda_stream) == L.MSS_EMPTY_CLASS


def test_idempotent_and_permutation_invariant(M):
    """size-independent properties at a BASELINE-sized image batch (4 x 1024 x 2048)."""
    n = 4 * 1024 * 2048
    g = torch.Generator(device="cuda").manual_seed(5)
    s = torch.randn(n, device="cuda", generator=g)
    lab = torch.where(torch.rand(n, device="cuda", generator=g) < 0.05, 1, 0).to(torch.uint8)
    lab[torch.rand(n, device="cuda", generator=g) > 0.95] = 255
    s = s + (lab == 1) * 1.5
    r1 = _tup(M.eval_ood_measure(s, lab))
    r2 = _tup(M.eval_ood_measure(s, lab))
    perm = torch.randperm(n, device="cuda", generator=g)
    r3 = _tup(M.eval_ood_measure(s[perm], lab[perm]))
    assert r1 == r2 == r3
    # monotone transform of the scores leaves all three metrics unchanged
    r4 = _tup(M.eval_ood_measure(s * 2.0, lab))
    assert r4 == r1
    assert r1 == c_oracle.eval_ood_measure(s.cpu().numpy(), lab.cpu().numpy())
