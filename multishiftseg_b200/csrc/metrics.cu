// (a9-a13) exact tie-aware AUROC / AP / FPR@95TPR from sorted (key, label) pairs.
//
// Integer stage (bit-exact by construction):
//   thresholds = ends of runs of equal keys (== np.where(np.diff(y_score)) + last index,
//   sklearn _ranking.py:916-920, metric.py:110-111);  tps[k] = #positives at positions <= end_k
//   (cumsum(y)[idx], _ranking.py:1034 / metric.py:114), fps[k] = 1 + end_k - tps[k].
// float64 tail (must reproduce numpy/sklearn rounding exactly, so: explicit __d*_rn intrinsics, no
// FMA contraction, and numpy's pairwise summation tree replayed leaf by leaf):
//   AUROC  roc_curve(drop_intermediate=True) + auc/trapezoid   _ranking.py:1331-1378, :53-116
//   AP     precision_recall_curve + step integral              _ranking.py:1160-1208, :243-260
//   FPR95  fpr_and_fdr_at_recall                               metric.py:116-127
// This translation unit is compiled with -fmad=false.
#include <vector>

#include "common.cuh"

namespace mss {

constexpr int CT_THREADS = 256;
constexpr int CT_IPT = 8;
constexpr int CT_TILE = CT_THREADS * CT_IPT;  // 2048

struct Pair64 {
    unsigned long long a, b;
};

// block-wide exclusive scan of one (a, b) pair per thread; returns exclusive prefix, total in `tot`
__device__ __forceinline__ uint2 block_exclusive_scan2(uint2 v, uint2 &tot) {
    __shared__ uint2 s_w[CT_THREADS / 32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint2 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned ta = __shfl_up_sync(0xffffffffu, inc.x, d), tb = __shfl_up_sync(0xffffffffu, inc.y, d);
        if (lane >= d) { inc.x += ta; inc.y += tb; }
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    uint2 base = make_uint2(0, 0), all = make_uint2(0, 0);
#pragma unroll
    for (int w = 0; w < CT_THREADS / 32; w++) {
        uint2 x = s_w[w];
        if (w < (int)warp) { base.x += x.x; base.y += x.y; }
        all.x += x.x; all.y += x.y;
    }
    tot = all;
    __syncthreads();
    return make_uint2(base.x + inc.x - v.x, base.y + inc.y - v.y);
}

// ---- sorted pairs -> (tps, fps) ---------------------------------------------------------------------
// phase A (SCATTER=false): per-tile (#run ends, #positives).  phase C (SCATTER=true): write counts.
template <bool SCATTER>
__global__ void __launch_bounds__(CT_THREADS)
runs_kernel(const uint32_t *__restrict__ keys, const uint8_t *__restrict__ labs, long long n,
            uint2 *__restrict__ tile_sums, const Pair64 *__restrict__ tile_excl, long long pos_before,
            long long idx_before, long long *__restrict__ tps, long long *__restrict__ fps) {
    const long long tile0 = (long long)blockIdx.x * CT_TILE;
    const long long i0 = tile0 + (long long)threadIdx.x * CT_IPT;
    uint32_t k[CT_IPT + 1];
    uint8_t l[CT_IPT];
#pragma unroll
    for (int j = 0; j < CT_IPT; j++) {
        k[j] = (i0 + j < n) ? __ldg(keys + i0 + j) : 0u;
        l[j] = (i0 + j < n) ? __ldg(labs + i0 + j) : (uint8_t)0;
    }
    k[CT_IPT] = (i0 + CT_IPT < n) ? __ldg(keys + i0 + CT_IPT) : 0u;
    unsigned ends = 0, npos = 0;
#pragma unroll
    for (int j = 0; j < CT_IPT; j++) {
        const long long i = i0 + j;
        if (i < n) {
            const bool end = (i == n - 1) || (k[j] != k[j + 1]);
            ends += end;
            npos += l[j];
        }
    }
    uint2 tot;
    uint2 ex = block_exclusive_scan2(make_uint2(ends, npos), tot);
    if (!SCATTER) {
        if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
        return;
    }
    const Pair64 te = tile_excl[blockIdx.x];
    unsigned long long kidx = te.a + ex.x;
    unsigned long long cp = te.b + ex.y;
#pragma unroll
    for (int j = 0; j < CT_IPT; j++) {
        const long long i = i0 + j;
        if (i < n) {
            cp += l[j];
            const bool end = (i == n - 1) || (k[j] != k[j + 1]);
            if (end) {
                const long long t = pos_before + (long long)cp;
                tps[kidx] = t;
                fps[kidx] = idx_before + i + 1 - t;
                kidx++;
            }
        }
    }
}

// exclusive scan of per-tile (a, b) sums by ONE block (tiles <= a few million); totals -> out[0..1]
__global__ void __launch_bounds__(1024)
tile_scan_kernel(const uint2 *__restrict__ sums, long long tiles, Pair64 *__restrict__ excl,
                 unsigned long long *__restrict__ totals) {
    __shared__ unsigned long long s_a[32], s_b[32];
    __shared__ unsigned long long s_carry[2];
    if (threadIdx.x == 0) { s_carry[0] = 0; s_carry[1] = 0; }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < tiles; base += 1024) {
        const long long t = base + threadIdx.x;
        uint2 v = (t < tiles) ? sums[t] : make_uint2(0, 0);
        unsigned long long a = v.x, b = v.y, ia = a, ib = b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long ta = __shfl_up_sync(0xffffffffu, ia, d), tb = __shfl_up_sync(0xffffffffu, ib, d);
            if (lane >= d) { ia += ta; ib += tb; }
        }
        if (lane == 31) { s_a[warp] = ia; s_b[warp] = ib; }
        __syncthreads();
        unsigned long long wa = 0, wb = 0, alla = 0, allb = 0;
        for (int w = 0; w < 32; w++) {
            if (w < (int)warp) { wa += s_a[w]; wb += s_b[w]; }
            alla += s_a[w]; allb += s_b[w];
        }
        const unsigned long long ca = s_carry[0], cb = s_carry[1];
        if (t < tiles) { excl[t].a = ca + wa + ia - a; excl[t].b = cb + wb + ib - b; }
        __syncthreads();
        if (threadIdx.x == 0) { s_carry[0] = ca + alla; s_carry[1] = cb + allb; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = s_carry[0]; totals[1] = s_carry[1]; }
}

// ---- ROC: drop collinear points (roc_curve drop_intermediate, _ranking.py:1338-1350) ----------------
__device__ __forceinline__ bool roc_keep(const long long *__restrict__ tps, const long long *__restrict__ fps,
                                         long long k, long long T) {
    if (T <= 2 || k == 0 || k == T - 1) return true;
    const long long d2f = fps[k + 1] - 2 * fps[k] + fps[k - 1];
    const long long d2t = tps[k + 1] - 2 * tps[k] + tps[k - 1];
    return d2f != 0 || d2t != 0;
}

template <bool SCATTER>
__global__ void __launch_bounds__(CT_THREADS)
roc_compact_kernel(const long long *__restrict__ tps, const long long *__restrict__ fps, long long T,
                   uint2 *__restrict__ tile_sums, const Pair64 *__restrict__ tile_excl,
                   long long *__restrict__ tps_k, long long *__restrict__ fps_k) {
    const long long i0 = (long long)blockIdx.x * CT_TILE + (long long)threadIdx.x * CT_IPT;
    unsigned keep = 0, cnt = 0;
#pragma unroll
    for (int j = 0; j < CT_IPT; j++) {
        const long long k = i0 + j;
        if (k < T && roc_keep(tps, fps, k, T)) { keep |= 1u << j; cnt++; }
    }
    uint2 tot;
    uint2 ex = block_exclusive_scan2(make_uint2(cnt, 0), tot);
    if (!SCATTER) {
        if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
        return;
    }
    unsigned long long o = tile_excl[blockIdx.x].a + ex.x;
#pragma unroll
    for (int j = 0; j < CT_IPT; j++) {
        if ((keep >> j) & 1u) {
            tps_k[o] = tps[i0 + j];
            fps_k[o] = fps[i0 + j];
            o++;
        }
    }
}

// ---- numpy pairwise-sum leaves ---------------------------------------------------------------------
// AP term j (reversed order, j = 0..T-1, k = T-1-j):  (rec[k-1] - rec[k]) * prec[k],  rec[-1] := 0
//   == diff(recall)[j] * precision[j] of the reversed, (1,0)-appended arrays.
struct ApTerm {
    const long long *tps, *fps;
    long long T;
    double P;
    __device__ __forceinline__ double operator()(long long j) const {
        const long long k = T - 1 - j;
        const double t = (double)tps[k], f = (double)fps[k];
        const double rec_k = __ddiv_rn(t, P);
        const double rec_prev = (k > 0) ? __ddiv_rn((double)tps[k - 1], P) : 0.0;
        const double prec = __ddiv_rn(t, __dadd_rn(t, f));
        return __dmul_rn(__dsub_rn(rec_prev, rec_k), prec);
    }
};
// ROC term j (j = 0..T'-1) over points p_0 = (0,0), p_{j+1} = (fps_k[j]/N, tps_k[j]/P):
//   ((fpr[j+1] - fpr[j]) * (tpr[j+1] + tpr[j])) / 2.0        scipy trapezoid
struct RocTerm {
    const long long *tps_k, *fps_k;
    double P, N;
    __device__ __forceinline__ double operator()(long long j) const {
        const double f1 = __ddiv_rn((double)fps_k[j], N), t1 = __ddiv_rn((double)tps_k[j], P);
        const double f0 = (j > 0) ? __ddiv_rn((double)fps_k[j - 1], N) : __ddiv_rn(0.0, N);
        const double t0 = (j > 0) ? __ddiv_rn((double)tps_k[j - 1], P) : __ddiv_rn(0.0, P);
        return __ddiv_rn(__dmul_rn(__dsub_rn(f1, f0), __dadd_rn(t1, t0)), 2.0);
    }
};

// 8 lanes per leaf = numpy's 8 interleaved accumulators; leaf_start has n_leaves+1 entries.
template <typename Term>
__global__ void __launch_bounds__(256)
leaf_sum_kernel(Term term, const long long *__restrict__ leaf_start, long long n_leaves,
                double *__restrict__ leaf_sum) {
    const long long leaf = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const unsigned sub = threadIdx.x & 7;
    const bool live = leaf < n_leaves;
    const long long s = live ? leaf_start[leaf] : 0;
    const long long m = live ? leaf_start[leaf + 1] - s : 0;
    double res;
    if (m < 8) {
        // whole array shorter than 8 terms: sequential, starting from -0.0
        res = -0.0;
        if (sub == 0)
            for (long long i = 0; i < m; i++) res = __dadd_rn(res, term(s + i));
    } else {
        const long long body = m - (m & 7);
        double r = term(s + sub);
        for (long long i = 8; i < body; i += 8) r = __dadd_rn(r, term(s + i + sub));
        // ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)) -- fp addition is commutative, so the butterfly is exact
        r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
        res = r;
        if (sub == 0)
            for (long long i = body; i < m; i++) res = __dadd_rn(res, term(s + i));
    }
    if (live && sub == 0) leaf_sum[leaf] = res;
}

// ---- FPR@95: argmin_k |tps[k]/P - 0.95| over k <= first k with tps[k]==P, ties -> largest k ----------
struct Best {
    double d;
    long long k;
};
__device__ __forceinline__ Best better(Best x, Best y) {
    if (y.d < x.d || (y.d == x.d && y.k > x.k)) return y;
    return x;
}

__global__ void __launch_bounds__(256)
fpr_partial_kernel(const long long *__restrict__ tps, long long T, double P, double recall_level,
                   Best *__restrict__ partial) {
    Best b{INFINITY, -1};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < T; k += stride) {
        const bool in_range = (k == 0) || ((double)tps[k - 1] != P);   // k <= searchsorted(tps, tps[-1])
        if (in_range) {
            const double d = fabs(__dsub_rn(__ddiv_rn((double)tps[k], P), recall_level));
            b = better(b, Best{d, k});
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        Best o{__shfl_xor_sync(0xffffffffu, b.d, s), __shfl_xor_sync(0xffffffffu, b.k, s)};
        b = better(b, o);
    }
    __shared__ Best sb[8];
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) b = better(b, sb[w]);
        partial[blockIdx.x] = b;
    }
}

__global__ void __launch_bounds__(256)
fpr_final_kernel(const Best *__restrict__ partial, int n, const long long *__restrict__ fps, double N,
                 double *__restrict__ out, long long *__restrict__ kout) {
    Best b{INFINITY, -1};
    for (int i = threadIdx.x; i < n; i += 256) b = better(b, partial[i]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        Best o{__shfl_xor_sync(0xffffffffu, b.d, s), __shfl_xor_sync(0xffffffffu, b.k, s)};
        b = better(b, o);
    }
    __shared__ Best sb[8];
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) b = better(b, sb[w]);
        out[0] = __ddiv_rn((double)fps[b.k], N);
        kout[0] = b.k;
    }
}

// ---- host side of the pairwise tree ----------------------------------------------------------------
static void pairwise_leaves(long long n, std::vector<long long> &starts) {
    starts.clear();
    // iterative DFS, left first; a leaf is a segment of <= 128 terms (numpy PW_BLOCKSIZE)
    std::vector<std::pair<long long, long long>> stack;
    stack.emplace_back(0, n);
    while (!stack.empty()) {
        auto [s, m] = stack.back();
        stack.pop_back();
        if (m <= 128) {
            starts.push_back(s);
        } else {
            long long n2 = m / 2;
            n2 -= n2 % 8;
            stack.emplace_back(s + n2, m - n2);
            stack.emplace_back(s, n2);
        }
    }
    starts.push_back(n);
}

static double pairwise_combine(const double *leaf, size_t &next, long long m) {
    if (m <= 128) return leaf[next++];
    long long n2 = m / 2;
    n2 -= n2 % 8;
    const double a = pairwise_combine(leaf, next, n2);
    const double b = pairwise_combine(leaf, next, m - n2);
    volatile double r = a + b;     // one rounded float64 add (no excess precision on any host)
    return r;
}

static size_t ct_tiles(int64_t n) { return (size_t)((n + CT_TILE - 1) / CT_TILE); }

}  // namespace mss

using namespace mss;

extern "C" size_t mss_counts_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    return ct_tiles(n) * (sizeof(uint2) + sizeof(Pair64)) + 1024;
}

extern "C" int mss_counts_from_sorted(const uint32_t *keys, const uint8_t *labs, int64_t n, int64_t pos_before,
                                      int64_t idx_before, int64_t *tps, int64_t *fps, int64_t *T_host,
                                      int64_t pn_host[2], void *workspace, size_t workspace_bytes, void *stream) {
    MSS_REQUIRE(n >= 0 && T_host && pn_host, "mss_counts_from_sorted: bad arguments");
    *T_host = 0; pn_host[0] = pn_host[1] = 0;
    if (n == 0) return MSS_OK;
    MSS_REQUIRE(keys && labs && tps && fps && workspace, "mss_counts_from_sorted: null pointer");
    const size_t tiles = ct_tiles(n);
    MSS_REQUIRE(tiles < (1ull << 31), "mss_counts_from_sorted: n too large");
    Carver c(workspace, workspace_bytes);
    uint2 *sums = c.take<uint2>(tiles);
    Pair64 *excl = c.take<Pair64>(tiles);
    unsigned long long *totals = c.take<unsigned long long>(2);
    if (!c.ok()) {
        set_error("mss_counts_from_sorted: workspace too small (%zu < %zu)", workspace_bytes, mss_counts_workspace_bytes(n));
        return MSS_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    runs_kernel<false><<<(unsigned)tiles, CT_THREADS, 0, st>>>(keys, labs, n, sums, nullptr, 0, 0, nullptr, nullptr);
    MSS_CHECK_LAUNCH();
    tile_scan_kernel<<<1, 1024, 0, st>>>(sums, (long long)tiles, excl, totals);
    MSS_CHECK_LAUNCH();
    runs_kernel<true><<<(unsigned)tiles, CT_THREADS, 0, st>>>(keys, labs, n, nullptr, excl, pos_before, idx_before,
                                                             (long long *)tps, (long long *)fps);
    MSS_CHECK_LAUNCH();
    unsigned long long h[2];
    MSS_CHECK_CUDA(cudaMemcpyAsync(h, totals, sizeof(h), cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    *T_host = (int64_t)h[0];
    pn_host[0] = (int64_t)h[1];
    pn_host[1] = n - (int64_t)h[1];
    return MSS_OK;
}

extern "C" size_t mss_tail_workspace_bytes(int64_t T) {
    if (T < 0) T = 0;
    const size_t leaves = (size_t)T / 64 + 2;
    return 2 * align_up((size_t)T * 8, 256)                      /* tps_k, fps_k */
           + ct_tiles(T) * (sizeof(uint2) + sizeof(Pair64))      /* compaction tiles */
           + 2 * align_up((leaves + 1) * 8, 256)                 /* leaf_start (AP, ROC) */
           + 2 * align_up(leaves * 8, 256)                       /* leaf sums */
           + 1024 * sizeof(Best) + 4096;
}

extern "C" int mss_metrics_tail(const int64_t *tps_, const int64_t *fps_, int64_t T, double recall_level,
                                void *workspace, size_t workspace_bytes, double out_host[3], int64_t *T_roc_host,
                                void *stream) {
    MSS_REQUIRE(tps_ && fps_ && T >= 1 && workspace && out_host, "mss_metrics_tail: bad arguments");
    const long long *tps = (const long long *)tps_, *fps = (const long long *)fps_;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t tiles = ct_tiles(T);
    const size_t max_leaves = (size_t)T / 64 + 2;
    Carver c(workspace, workspace_bytes);
    long long *tps_k = c.take<long long>((size_t)T);
    long long *fps_k = c.take<long long>((size_t)T);
    uint2 *sums = c.take<uint2>(tiles);
    Pair64 *excl = c.take<Pair64>(tiles);
    long long *leaf_ap = c.take<long long>(max_leaves + 1);
    long long *leaf_roc = c.take<long long>(max_leaves + 1);
    double *sum_ap = c.take<double>(max_leaves);
    double *sum_roc = c.take<double>(max_leaves);
    Best *partial = c.take<Best>(1024);
    unsigned long long *totals = c.take<unsigned long long>(2);
    double *fpr_out = c.take<double>(1);
    long long *fpr_k = c.take<long long>(1);
    if (!c.ok()) {
        set_error("mss_metrics_tail: workspace too small (%zu < %zu)", workspace_bytes, mss_tail_workspace_bytes(T));
        return MSS_ERR_WORKSPACE;
    }
    // P, N = last entries
    long long PN[2];
    MSS_CHECK_CUDA(cudaMemcpyAsync(&PN[0], tps + (T - 1), 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(&PN[1], fps + (T - 1), 8, cudaMemcpyDeviceToHost, st));

    // ROC compaction (count, scan, scatter)
    roc_compact_kernel<false><<<(unsigned)tiles, CT_THREADS, 0, st>>>(tps, fps, T, sums, nullptr, nullptr, nullptr);
    MSS_CHECK_LAUNCH();
    tile_scan_kernel<<<1, 1024, 0, st>>>(sums, (long long)tiles, excl, totals);
    MSS_CHECK_LAUNCH();
    roc_compact_kernel<true><<<(unsigned)tiles, CT_THREADS, 0, st>>>(tps, fps, T, nullptr, excl, tps_k, fps_k);
    MSS_CHECK_LAUNCH();
    unsigned long long h_tot[2];
    MSS_CHECK_CUDA(cudaMemcpyAsync(h_tot, totals, sizeof(h_tot), cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    const long long T_roc = (long long)h_tot[0];
    const double P = (double)PN[0], N = (double)PN[1];
    if (T_roc_host) *T_roc_host = T_roc;
    if (PN[0] <= 0 || PN[1] <= 0) {
        set_error("mss_metrics_tail: P=%lld N=%lld (a class is empty)", PN[0], PN[1]);
        return MSS_EMPTY_CLASS;
    }

    // leaf tables (host) -> device; leaf sums (device) -> host; tree combine (host)
    std::vector<long long> l_ap, l_roc;
    pairwise_leaves(T, l_ap);
    pairwise_leaves(T_roc, l_roc);
    const long long n_ap = (long long)l_ap.size() - 1, n_roc = (long long)l_roc.size() - 1;
    if ((size_t)n_ap > max_leaves || (size_t)n_roc > max_leaves) {
        set_error("mss_metrics_tail: internal leaf-count bound exceeded");
        return MSS_ERR_WORKSPACE;
    }
    MSS_CHECK_CUDA(cudaMemcpyAsync(leaf_ap, l_ap.data(), l_ap.size() * 8, cudaMemcpyHostToDevice, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(leaf_roc, l_roc.data(), l_roc.size() * 8, cudaMemcpyHostToDevice, st));
    leaf_sum_kernel<ApTerm><<<(unsigned)((n_ap * 8 + 255) / 256), 256, 0, st>>>(ApTerm{tps, fps, T, P}, leaf_ap, n_ap, sum_ap);
    MSS_CHECK_LAUNCH();
    leaf_sum_kernel<RocTerm><<<(unsigned)((n_roc * 8 + 255) / 256), 256, 0, st>>>(RocTerm{tps_k, fps_k, P, N}, leaf_roc, n_roc, sum_roc);
    MSS_CHECK_LAUNCH();
    int fgrid = (int)std::min<long long>((T + 255) / 256, 1024);
    fpr_partial_kernel<<<fgrid, 256, 0, st>>>(tps, T, P, recall_level, partial);
    MSS_CHECK_LAUNCH();
    fpr_final_kernel<<<1, 256, 0, st>>>(partial, fgrid, fps, N, fpr_out, fpr_k);
    MSS_CHECK_LAUNCH();
    std::vector<double> h_ap((size_t)n_ap), h_roc((size_t)n_roc);
    double h_fpr = 0.0;
    MSS_CHECK_CUDA(cudaMemcpyAsync(h_ap.data(), sum_ap, (size_t)n_ap * 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(h_roc.data(), sum_roc, (size_t)n_roc * 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(&h_fpr, fpr_out, 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    size_t nx = 0;
    const double auroc = pairwise_combine(h_roc.data(), nx, T_roc);
    nx = 0;
    const double ap_sum = pairwise_combine(h_ap.data(), nx, T);
    const double ap = -ap_sum;
    out_host[0] = auroc;                  // auc(): direction == 1 because fpr is non-decreasing
    out_host[1] = ap > 0.0 ? ap : 0.0;    // max(0.0, -sum(...))
    out_host[2] = h_fpr;
    return MSS_OK;
}

// ---- composites ------------------------------------------------------------------------------------
static size_t one_shot_layout(int64_t n, size_t off[6]) {
    // [keys n*4][labs n][state][sort ws][tps n*8][fps n*8][counts ws][tail ws]
    size_t o = 0;
    off[0] = o; o += align_up((size_t)n * 4, 256);
    off[1] = o; o += align_up((size_t)n, 256);
    off[2] = o; o += 256;
    off[3] = o; o += align_up(mss_sort_pairs_workspace_bytes(n), 256);
    off[4] = o; o += 2 * align_up((size_t)n * 8, 256);
    off[5] = o; o += align_up(mss_counts_workspace_bytes(n), 256) + align_up(mss_tail_workspace_bytes(n), 256);
    return o;
}

extern "C" size_t mss_ood_metrics_workspace_bytes(int64_t n) {
    size_t off[6];
    return one_shot_layout(n < 0 ? 0 : n, off) + 256;
}

static int metrics_from_pairs(uint32_t *keys, uint8_t *labs, int64_t m, char *ws_sort, size_t sort_bytes,
                              int64_t *tps, int64_t *fps, char *ws_rest, size_t rest_bytes, double out_host[3],
                              int64_t counts_host[4], void *stream) {
    int rc = mss_sort_pairs(keys, labs, m, ws_sort, sort_bytes, stream);
    if (rc) return rc;
    int64_t T = 0, pn[2];
    const size_t cbytes = align_up(mss_counts_workspace_bytes(m), 256);
    rc = mss_counts_from_sorted(keys, labs, m, 0, 0, tps, fps, &T, pn, ws_rest, cbytes, stream);
    if (rc) return rc;
    int64_t T_roc = 0;
    rc = mss_metrics_tail(tps, fps, T, 0.95, ws_rest + cbytes, rest_bytes - cbytes, out_host, &T_roc, stream);
    if (rc) return rc;
    if (counts_host) { counts_host[0] = pn[0]; counts_host[1] = pn[1]; counts_host[2] = T; counts_host[3] = T_roc; }
    return MSS_OK;
}

static int check_state(const int64_t st[4]) {
    if (st[2]) { set_error("Input contains NaN."); return MSS_ERR_NAN; }
    if (st[3]) { set_error("Input contains infinity or a value too large for dtype('float32')."); return MSS_ERR_INF; }
    if (st[1] == 0 || st[1] == st[0]) return MSS_EMPTY_CLASS;
    return MSS_OK;
}

extern "C" int mss_ood_metrics(const float *scores, const void *labels, int label_dtype, int64_t n, int64_t id_in,
                               int64_t id_out, void *workspace, size_t workspace_bytes, double out_host[3],
                               int64_t counts_host[4], void *stream) {
    MSS_REQUIRE(n >= 0 && out_host, "mss_ood_metrics: bad arguments");
    if (n == 0) return MSS_EMPTY_CLASS;
    MSS_REQUIRE(scores && labels && workspace, "mss_ood_metrics: null pointer");
    size_t off[6];
    const size_t need = one_shot_layout(n, off);
    char *ws = (char *)align_up((size_t)(uintptr_t)workspace, 256);
    if ((size_t)(ws - (char *)workspace) + need > workspace_bytes) {
        set_error("mss_ood_metrics: workspace too small (%zu < %zu)", workspace_bytes, mss_ood_metrics_workspace_bytes(n));
        return MSS_ERR_WORKSPACE;
    }
    mss_eval_buffers ev{(uint32_t *)(ws + off[0]), (uint8_t *)(ws + off[1]), ws + off[2], n};
    int rc = mss_eval_reset(&ev, stream);
    if (rc) return rc;
    rc = mss_eval_append(scores, labels, label_dtype, n, id_in, id_out, &ev, stream);
    if (rc) return rc;
    int64_t st[4];
    rc = mss_eval_state_host(&ev, st, stream);
    if (rc) return rc;
    // sklearn validates before anything else, the reference checks emptiness first (metric.py:176)
    if (st[1] == 0 || st[1] == st[0]) return MSS_EMPTY_CLASS;
    rc = check_state(st);
    if (rc) return rc;
    return metrics_from_pairs(ev.keys, ev.labs, st[0], ws + off[3], off[4] - off[3], (int64_t *)(ws + off[4]),
                              (int64_t *)(ws + off[4] + align_up((size_t)n * 8, 256)), ws + off[5], need - off[5],
                              out_host, counts_host, stream);
}

extern "C" int mss_ood_metrics_from_eval(const mss_eval_buffers *ev, void *workspace, size_t workspace_bytes,
                                         double out_host[3], int64_t counts_host[4], void *stream) {
    MSS_REQUIRE(ev && ev->keys && ev->labs && ev->state && workspace && out_host, "mss_ood_metrics_from_eval: null pointer");
    int64_t st[4];
    int rc = mss_eval_state_host(ev, st, stream);
    if (rc) return rc;
    if (st[1] == 0 || st[1] == st[0]) return MSS_EMPTY_CLASS;
    rc = check_state(st);
    if (rc) return rc;
    const int64_t m = st[0];
    char *ws = (char *)align_up((size_t)(uintptr_t)workspace, 256);
    const size_t lead = (size_t)(ws - (char *)workspace);
    const size_t sort_b = align_up(mss_sort_pairs_workspace_bytes(m), 256);
    const size_t cnt_b = 2 * align_up((size_t)m * 8, 256);
    const size_t rest_b = align_up(mss_counts_workspace_bytes(m), 256) + align_up(mss_tail_workspace_bytes(m), 256);
    if (lead + sort_b + cnt_b + rest_b > workspace_bytes) {
        set_error("mss_ood_metrics_from_eval: workspace too small (%zu < %zu)", workspace_bytes, lead + sort_b + cnt_b + rest_b);
        return MSS_ERR_WORKSPACE;
    }
    return metrics_from_pairs(ev->keys, ev->labs, m, ws, sort_b, (int64_t *)(ws + sort_b),
                              (int64_t *)(ws + sort_b + align_up((size_t)m * 8, 256)), ws + sort_b + cnt_b, rest_b,
                              out_host, counts_host, stream);
}
