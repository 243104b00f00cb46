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
#include <map>
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
    static_assert(CT_IPT == 8, "vector path below loads 2 x uint4 keys + 1 x uint2 labels");
    if (i0 + CT_IPT <= n && (((uintptr_t)keys & 15) | ((uintptr_t)labs & 7)) == 0) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(keys + i0));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(keys + i0 + 4));
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(labs + i0));
        k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            l[j] = (uint8_t)(v.x >> (8 * j));
            l[4 + j] = (uint8_t)(v.y >> (8 * j));
        }
    } else {
#pragma unroll
        for (int j = 0; j < CT_IPT; j++) {
            k[j] = (i0 + j < n) ? __ldg(keys + i0 + j) : 0u;
            l[j] = (i0 + j < n) ? __ldg(labs + i0 + j) : (uint8_t)0;
        }
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
    // stage the tile's thresholds in shared memory (thread-order == output order), then write them out coalesced
    __shared__ long long s_t[CT_TILE], s_f[CT_TILE];
    const Pair64 te = tile_excl[blockIdx.x];
    unsigned slot = ex.x;
    unsigned long long cp = te.b + ex.y;
#pragma unroll
    for (int j = 0; j < CT_IPT; j++) {
        const long long i = i0 + j;
        if (i < n) {
            cp += l[j];
            const bool end = (i == n - 1) || (k[j] != k[j + 1]);
            if (end) {
                const long long t = pos_before + (long long)cp;
                s_t[slot] = t;
                s_f[slot] = idx_before + i + 1 - t;
                slot++;
            }
        }
    }
    __syncthreads();
    for (unsigned q = threadIdx.x; q < tot.x; q += CT_THREADS) {
        tps[te.a + q] = s_t[q];
        fps[te.a + q] = s_f[q];
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
// ---- FPR@95: argmin_k |tps[k]/P - 0.95| over k <= first k with tps[k]==P, ties -> largest k -----------
//      (fpr_and_fdr_at_recall, metric.py:116-127; fused into the counting pass below: same data window)
struct Best {
    double d;
    long long k;
};
__device__ __forceinline__ Best better(Best x, Best y) {
    if (y.d < x.d || (y.d == x.d && y.k > x.k)) return y;
    return x;
}
__device__ __forceinline__ Best block_best(Best b) {      // valid in thread 0
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        Best o{__shfl_xor_sync(0xffffffffu, b.d, s), __shfl_xor_sync(0xffffffffu, b.k, s)};
        b = better(b, o);
    }
    __shared__ Best sb[CT_THREADS / 32];
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < CT_THREADS / 32; w++) b = better(b, sb[w]);
    return b;
}

// One thread owns thresholds [i0, i0+8); it needs (tps, fps) at i0-1 .. i0+8: 10 + 10 loads instead of the
// 48 a per-point second difference would issue.  SCATTER=false: per-tile kept count (+ the FPR95 candidate
// of the tile); SCATTER=true: write the kept points.
template <bool SCATTER>
__global__ void __launch_bounds__(CT_THREADS)
roc_compact_kernel(const long long *__restrict__ tps, const long long *__restrict__ fps, long long T,
                   double recall_level, uint2 *__restrict__ tile_sums, Best *__restrict__ tile_best,
                   const Pair64 *__restrict__ tile_excl, long long *__restrict__ tps_k,
                   long long *__restrict__ fps_k) {
    const long long i0 = (long long)blockIdx.x * CT_TILE + (long long)threadIdx.x * CT_IPT;
    long long t[CT_IPT + 2], f[CT_IPT + 2];                  // window index w <-> threshold i0 - 1 + w
    if (i0 + CT_IPT <= T && ((((uintptr_t)tps) | ((uintptr_t)fps)) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < CT_IPT; j += 2) {
            const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(tps + i0 + j));
            const longlong2 c = __ldg(reinterpret_cast<const longlong2 *>(fps + i0 + j));
            t[j + 1] = a.x; t[j + 2] = a.y;
            f[j + 1] = c.x; f[j + 2] = c.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < CT_IPT; j++) {
            const bool in = i0 + j < T;
            t[j + 1] = in ? __ldg(tps + i0 + j) : 0;
            f[j + 1] = in ? __ldg(fps + i0 + j) : 0;
        }
    }
    t[0] = (i0 > 0 && i0 - 1 < T) ? __ldg(tps + i0 - 1) : 0;
    f[0] = (i0 > 0 && i0 - 1 < T) ? __ldg(fps + i0 - 1) : 0;
    t[CT_IPT + 1] = (i0 + CT_IPT < T) ? __ldg(tps + i0 + CT_IPT) : 0;
    f[CT_IPT + 1] = (i0 + CT_IPT < T) ? __ldg(fps + i0 + CT_IPT) : 0;

    unsigned keep = 0, cnt = 0;
#pragma unroll
    for (int j = 0; j < CT_IPT; j++) {
        const long long k = i0 + j;
        if (k < T) {
            bool kp = true;
            if (T > 2 && k != 0 && k != T - 1) {
                const long long d2f = f[j + 2] - 2 * f[j + 1] + f[j];
                const long long d2t = t[j + 2] - 2 * t[j + 1] + t[j];
                kp = d2f != 0 || d2t != 0;
            }
            if (kp) { keep |= 1u << j; cnt++; }
        }
    }
    uint2 tot;
    uint2 ex = block_exclusive_scan2(make_uint2(cnt, 0), tot);
    if (!SCATTER) {
        if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
        // FPR95 candidate of this tile
        const double P = (double)__ldg(tps + T - 1);
        Best b{INFINITY, -1};
#pragma unroll
        for (int j = 0; j < CT_IPT; j++) {
            const long long k = i0 + j;
            if (k < T && (k == 0 || (double)t[j] != P)) {     // k <= searchsorted(tps, tps[-1])
                const double d = fabs(__dsub_rn(__ddiv_rn((double)t[j + 1], P), recall_level));
                b = better(b, Best{d, k});
            }
        }
        b = block_best(b);
        if (threadIdx.x == 0) tile_best[blockIdx.x] = b;
        return;
    }
    unsigned long long o = tile_excl[blockIdx.x].a + ex.x;
#pragma unroll
    for (int j = 0; j < CT_IPT; j++) {
        if ((keep >> j) & 1u) {
            tps_k[o] = t[j + 1];
            fps_k[o] = f[j + 1];
            o++;
        }
    }
}

// reduce n candidates to gridDim.x candidates (grid-stride)
__global__ void __launch_bounds__(CT_THREADS)
fpr_reduce_kernel(const Best *__restrict__ in, long long n, Best *__restrict__ out) {
    Best b{INFINITY, -1};
    for (long long i = (long long)blockIdx.x * CT_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * CT_THREADS)
        b = better(b, in[i]);
    b = block_best(b);
    if (threadIdx.x == 0) out[blockIdx.x] = b;
}

__global__ void __launch_bounds__(CT_THREADS)
fpr_final_kernel(const Best *__restrict__ partial, int n, const long long *__restrict__ fps, long long T,
                 double *__restrict__ out, long long *__restrict__ kout) {
    const double N = (double)fps[T - 1];
    Best b{INFINITY, -1};
    for (int i = threadIdx.x; i < n; i += CT_THREADS) b = better(b, partial[i]);
    b = block_best(b);
    if (threadIdx.x == 0) {
        out[0] = (b.k >= 0) ? __ddiv_rn((double)fps[b.k], N) : nan("");   // k < 0 only when P == 0 (caller reports it)
        kout[0] = b.k;
    }
}

// ---- numpy pairwise-sum leaves ---------------------------------------------------------------------
// AP term j (reversed order, j = 0..T-1, k = T-1-j):  (rec[k-1] - rec[k]) * prec[k],  rec[-1] := 0
//   == diff(recall)[j] * precision[j] of the reversed, (1,0)-appended arrays.
struct ApTerm {
    const long long *tps, *fps;
    long long T;
    __device__ __forceinline__ double operator()(long long j) const {
        const double P = (double)tps[T - 1];
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
    const long long *p_P, *p_N;      // device: tps[T-1], fps[T-1]
    __device__ __forceinline__ double operator()(long long j) const {
        const double P = (double)*p_P, N = (double)*p_N;
        const double f1 = __ddiv_rn((double)fps_k[j], N), t1 = __ddiv_rn((double)tps_k[j], P);
        const double f0 = (j > 0) ? __ddiv_rn((double)fps_k[j - 1], N) : __ddiv_rn(0.0, N);
        const double t0 = (j > 0) ? __ddiv_rn((double)tps_k[j - 1], P) : __ddiv_rn(0.0, P);
        return __ddiv_rn(__dmul_rn(__dsub_rn(f1, f0), __dadd_rn(t1, t0)), 2.0);
    }
};

// ---- numpy pairwise tree, addressed without materialising it ---------------------------------------
// np.add.reduce over float64 (DOUBLE_pairwise_sum): a segment of m > 128 terms is split at
// n2 = m/2 - (m/2 % 8) into [0,n2) + [n2,m); segments of <= 128 terms are leaves.  The tree depends on
// n only, and the subtree sizes that occur in it are few (27 distinct sizes for n = 4 194 304 000), so
// the host builds a table size -> #leaves and the device finds leaf i by descending from the root.
constexpr int PW_MAX_SIZES = 256;
constexpr int PW_MAX_FRONT = 16384;

__host__ __device__ __forceinline__ long long pw_split(long long m) {
    const long long n2 = m / 2;
    return n2 - (n2 % 8);
}

struct PwTree {
    const long long *size;     // ascending subtree sizes > 128 that occur in the tree of n terms
    const long long *leaves;   // number of leaves of a subtree of that size
    int n_sizes;
    long long n;               // terms
    long long n_leaves;
    __host__ __device__ __forceinline__ long long leaves_of(long long m) const {
        if (m <= 128) return 1;
        int lo = 0, hi = n_sizes - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (size[mid] < m) lo = mid + 1; else hi = mid;
        }
        return leaves[lo];
    }
    // leaf index -> [start, start + len)
    __host__ __device__ __forceinline__ void leaf_bounds(long long leaf, long long &start, long long &len) const {
        long long s = 0, m = n;
        while (m > 128) {
            const long long n2 = pw_split(m);
            const long long l = leaves_of(n2);
            if (leaf < l) m = n2;
            else { leaf -= l; s += n2; m -= n2; }
        }
        start = s;
        len = m;
    }
};

// one thread per leaf: start offsets of all leaves (leaf_start[n_leaves] = n)
__global__ void __launch_bounds__(256)
leaf_bounds_kernel(PwTree tree, long long *__restrict__ leaf_start) {
    const long long leaf = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf > tree.n_leaves) return;
    long long s = tree.n, m = 0;
    if (leaf < tree.n_leaves) tree.leaf_bounds(leaf, s, m);
    leaf_start[leaf] = s;
}

// 8 lanes per leaf = numpy's 8 interleaved accumulators
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

// The tree above the leaves.  The host cuts it at "frontier" nodes (the first node on each root path
// with <= F terms, at most PW_MAX_FRONT of them); one thread combines the leaf sums under one frontier
// node in numpy's order (left subtree, right subtree, one rounded add), the host combines the frontier.
struct PwFrontNode {
    long long m;            // terms under the node
    long long first_leaf;   // index of its first leaf
};

// numpy's recursion over the leaf sums under one node of m terms (left subtree, right subtree, one add)
__host__ __device__ inline double pw_subtree_combine(long long m, const double *leaf_sum, long long next) {
    long long sz[48];
    double acc[48];
    unsigned char st[48];
    int sp = 0;
    sz[0] = m;
    st[0] = 0;
    for (;;) {
        if (sz[sp] > 128) {                       // descend left
            st[sp] = 1;
            sz[sp + 1] = pw_split(sz[sp]);
            st[sp + 1] = 0;
            sp++;
            continue;
        }
        double val = leaf_sum[next++];
        sp--;
        for (;;) {                                // return `val` to the parent frame
            if (sp < 0) return val;
            if (st[sp] == 1) {                    // left done: keep it, descend right
                acc[sp] = val;
                st[sp] = 2;
                sz[sp + 1] = sz[sp] - pw_split(sz[sp]);
                st[sp + 1] = 0;
                sp++;
                break;
            }
#ifdef __CUDA_ARCH__
            val = __dadd_rn(acc[sp], val);        // both done
#else
            { volatile double r = acc[sp] + val; val = r; }
#endif
            sp--;
        }
    }
}

__global__ void __launch_bounds__(128)
subtree_combine_kernel(const PwFrontNode *__restrict__ nodes, int n_nodes, const double *__restrict__ leaf_sum,
                       double *__restrict__ node_sum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    node_sum[i] = pw_subtree_combine(nodes[i].m, leaf_sum, nodes[i].first_leaf);
}

// ---- host side of the pairwise tree ----------------------------------------------------------------
struct PwPlan {
    std::vector<long long> size, leaves;          // table (ascending sizes > 128)
    std::vector<PwFrontNode> front;               // frontier nodes, left to right
    long long n = 0, n_leaves = 0, F = 0;
};

static long long pw_count_leaves(long long m, std::map<long long, long long> &memo) {
    if (m <= 128) return 1;
    auto it = memo.find(m);
    if (it != memo.end()) return it->second;
    const long long n2 = pw_split(m);
    const long long l = pw_count_leaves(n2, memo) + pw_count_leaves(m - n2, memo);
    memo[m] = l;
    return l;
}

static void pw_frontier(long long m, long long F, long long &leaf, std::map<long long, long long> &memo,
                        std::vector<PwFrontNode> &out) {
    if (m <= F || m <= 128) {
        out.push_back(PwFrontNode{m, leaf});
        leaf += pw_count_leaves(m, memo);
        return;
    }
    const long long n2 = pw_split(m);
    pw_frontier(n2, F, leaf, memo, out);
    pw_frontier(m - n2, F, leaf, memo, out);
}

static bool pw_plan(long long n, PwPlan &p) {
    std::map<long long, long long> memo;
    p.n = n;
    p.n_leaves = pw_count_leaves(n, memo);
    p.size.clear(); p.leaves.clear(); p.front.clear();
    for (auto &kv : memo) { p.size.push_back(kv.first); p.leaves.push_back(kv.second); }
    if (p.size.empty()) { p.size.push_back(129); p.leaves.push_back(2); }   // never looked up (n <= 128)
    // frontier granularity: nodes of <= F terms, F grown until there are at most ~PW_MAX_FRONT/2 nodes
    p.F = 2048;
    while (n / p.F > PW_MAX_FRONT / 4) p.F *= 2;
    long long leaf = 0;
    pw_frontier(n, p.F, leaf, memo, p.front);
    return (int)p.size.size() <= PW_MAX_SIZES && (int)p.front.size() <= PW_MAX_FRONT && leaf == p.n_leaves;
}

// combine the frontier sums exactly as the recursion above the frontier would
static double pw_combine_top(const double *node_sum, size_t &next, long long m, long long F) {
    if (m <= F || m <= 128) return node_sum[next++];
    const long long n2 = pw_split(m);
    const double a = pw_combine_top(node_sum, next, n2, F);
    const double b = pw_combine_top(node_sum, next, m - n2, F);
    volatile double r = a + b;     // one rounded float64 add (no excess precision on any host)
    return r;
}

struct PwDevice {
    long long *size, *leaves, *leaf_start;
    PwFrontNode *front;
    double *node_sum;
};

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

static size_t pw_device_bytes() {
    return 2 * align_up(PW_MAX_SIZES * 8, 256) + align_up(PW_MAX_FRONT * sizeof(PwFrontNode), 256) +
           align_up(PW_MAX_FRONT * 8, 256);
}

extern "C" size_t mss_tail_workspace_bytes(int64_t T) {
    if (T < 0) T = 0;
    const size_t leaves = (size_t)T / 64 + 2;
    return 2 * align_up((size_t)T * 8, 256)                      /* tps_k, fps_k */
           + ct_tiles(T) * (sizeof(uint2) + sizeof(Pair64) + sizeof(Best)) + 1024   /* compaction tiles, FPR95 candidates */
           + 2 * align_up(leaves * 8, 256)                       /* leaf sums (AP, ROC) */
           + 2 * align_up((leaves + 1) * 8, 256)                 /* leaf starts (AP, ROC) */
           + 2 * pw_device_bytes()                               /* tree tables + frontier (AP, ROC) */
           + 1024 * sizeof(Best) + 8192;
}

static PwDevice pw_carve(Carver &c, size_t max_leaves) {
    PwDevice d;
    d.leaf_start = c.take<long long>(max_leaves + 1);
    d.size = c.take<long long>(PW_MAX_SIZES);
    d.leaves = c.take<long long>(PW_MAX_SIZES);
    d.front = c.take<PwFrontNode>(PW_MAX_FRONT);
    d.node_sum = c.take<double>(PW_MAX_FRONT);
    return d;
}

// upload the plan, sum the leaves, combine below the frontier; frontier sums stay on the device
template <typename Term>
static int pw_launch(const PwPlan &p, const PwDevice &d, Term term, double *leaf_sum, cudaStream_t st) {
    MSS_CHECK_CUDA(cudaMemcpyAsync(d.size, p.size.data(), p.size.size() * 8, cudaMemcpyHostToDevice, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(d.leaves, p.leaves.data(), p.leaves.size() * 8, cudaMemcpyHostToDevice, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(d.front, p.front.data(), p.front.size() * sizeof(PwFrontNode), cudaMemcpyHostToDevice, st));
    PwTree tree{d.size, d.leaves, (int)p.size.size(), p.n, p.n_leaves};
    leaf_bounds_kernel<<<(unsigned)((p.n_leaves + 1 + 255) / 256), 256, 0, st>>>(tree, d.leaf_start);
    MSS_CHECK_LAUNCH();
    leaf_sum_kernel<Term><<<(unsigned)((p.n_leaves * 8 + 255) / 256), 256, 0, st>>>(term, d.leaf_start, p.n_leaves, leaf_sum);
    MSS_CHECK_LAUNCH();
    const int nf = (int)p.front.size();
    subtree_combine_kernel<<<(nf + 127) / 128, 128, 0, st>>>(d.front, nf, leaf_sum, d.node_sum);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
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
    Best *tile_best = c.take<Best>(tiles);
    double *sum_ap = c.take<double>(max_leaves);
    double *sum_roc = c.take<double>(max_leaves);
    PwDevice d_ap = pw_carve(c, max_leaves), d_roc = pw_carve(c, max_leaves);
    Best *partial = c.take<Best>(1024);
    unsigned long long *totals = c.take<unsigned long long>(2);
    double *fpr_out = c.take<double>(1);
    long long *fpr_k = c.take<long long>(1);
    if (!c.ok()) {
        set_error("mss_metrics_tail: workspace too small (%zu < %zu)", workspace_bytes, mss_tail_workspace_bytes(T));
        return MSS_ERR_WORKSPACE;
    }
    // P, N = last entries (the kernels read them from the device; the host copy is for the empty-class check)
    long long PN[2];
    MSS_CHECK_CUDA(cudaMemcpyAsync(&PN[0], tps + (T - 1), 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(&PN[1], fps + (T - 1), 8, cudaMemcpyDeviceToHost, st));

    // ROC compaction (count, scan, scatter)
    roc_compact_kernel<false><<<(unsigned)tiles, CT_THREADS, 0, st>>>(tps, fps, T, recall_level, sums, tile_best, nullptr,
                                                                      nullptr, nullptr);
    MSS_CHECK_LAUNCH();
    tile_scan_kernel<<<1, 1024, 0, st>>>(sums, (long long)tiles, excl, totals);
    MSS_CHECK_LAUNCH();
    roc_compact_kernel<true><<<(unsigned)tiles, CT_THREADS, 0, st>>>(tps, fps, T, recall_level, nullptr, nullptr, excl,
                                                                     tps_k, fps_k);
    MSS_CHECK_LAUNCH();
    // FPR95: tile candidates -> (<= 1024 candidates) -> result
    {
        const Best *cand = tile_best;
        int ncand = (int)tiles;
        if (tiles > 1024) {
            fpr_reduce_kernel<<<1024, CT_THREADS, 0, st>>>(tile_best, (long long)tiles, partial);
            MSS_CHECK_LAUNCH();
            cand = partial;
            ncand = 1024;
        }
        fpr_final_kernel<<<1, CT_THREADS, 0, st>>>(cand, ncand, fps, T, fpr_out, fpr_k);
        MSS_CHECK_LAUNCH();
    }
    unsigned long long h_tot[2];
    MSS_CHECK_CUDA(cudaMemcpyAsync(h_tot, totals, sizeof(h_tot), cudaMemcpyDeviceToHost, st));

    // AP and FPR95 need nothing from the host: enqueue them behind the compaction
    PwPlan p_ap, p_roc;
    if (!pw_plan(T, p_ap) || (size_t)p_ap.n_leaves > max_leaves) {
        set_error("mss_metrics_tail: internal pairwise-plan bound exceeded (T=%lld)", (long long)T);
        return MSS_ERR_WORKSPACE;
    }
    int rc = pw_launch(p_ap, d_ap, ApTerm{tps, fps, T}, sum_ap, st);
    if (rc) return rc;

    MSS_CHECK_CUDA(cudaStreamSynchronize(st));      // T_roc decides the shape of the ROC tree
    const long long T_roc = (long long)h_tot[0];
    if (T_roc_host) *T_roc_host = T_roc;
    if (PN[0] <= 0 || PN[1] <= 0) {
        set_error("mss_metrics_tail: P=%lld N=%lld (a class is empty)", PN[0], PN[1]);
        return MSS_EMPTY_CLASS;
    }
    if (!pw_plan(T_roc, p_roc) || (size_t)p_roc.n_leaves > max_leaves) {
        set_error("mss_metrics_tail: internal pairwise-plan bound exceeded (T_roc=%lld)", T_roc);
        return MSS_ERR_WORKSPACE;
    }
    rc = pw_launch(p_roc, d_roc, RocTerm{tps_k, fps_k, tps + (T - 1), fps + (T - 1)}, sum_roc, st);
    if (rc) return rc;

    std::vector<double> h_ap(p_ap.front.size()), h_roc(p_roc.front.size());
    double h_fpr = 0.0;
    MSS_CHECK_CUDA(cudaMemcpyAsync(h_ap.data(), d_ap.node_sum, h_ap.size() * 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(h_roc.data(), d_roc.node_sum, h_roc.size() * 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(&h_fpr, fpr_out, 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    size_t nx = 0;
    const double auroc = pw_combine_top(h_roc.data(), nx, T_roc, p_roc.F);
    nx = 0;
    const double ap_sum = pw_combine_top(h_ap.data(), nx, T, p_ap.F);
    const double ap = -ap_sum;
    out_host[0] = auroc;                  // auc(): direction == 1 because fpr is non-decreasing
    out_host[1] = ap > 0.0 ? ap : 0.0;    // max(0.0, -sum(...))
    out_host[2] = h_fpr;
    return MSS_OK;
}

/* test hook (host only): bounds of leaf `leaf` of numpy's pairwise tree over n terms, computed by the same
 * descent the device uses; *n_leaves_out = number of leaves.  Returns MSS_ERR_INVALID_ARG if out of range. */
extern "C" int mss_pairwise_leaf_bounds(int64_t n, int64_t leaf, int64_t *start_out, int64_t *len_out,
                                        int64_t *n_leaves_out) {
    MSS_REQUIRE(n >= 1 && start_out && len_out && n_leaves_out, "mss_pairwise_leaf_bounds: bad arguments");
    PwPlan p;
    MSS_REQUIRE(pw_plan(n, p), "mss_pairwise_leaf_bounds: plan bound exceeded");
    *n_leaves_out = p.n_leaves;
    MSS_REQUIRE(leaf >= 0 && leaf < p.n_leaves, "mss_pairwise_leaf_bounds: leaf out of range");
    PwTree tree{p.size.data(), p.leaves.data(), (int)p.size.size(), p.n, p.n_leaves};
    long long s, m;
    tree.leaf_bounds(leaf, s, m);
    *start_out = s;
    *len_out = m;
    return MSS_OK;
}

/* test hook (host only): sum n float64 terms through the SAME plan / descent / subtree-combine / top-combine
 * code the device path uses (the leaves themselves are summed here with numpy's 8-accumulator loop). */
extern "C" int mss_pairwise_sum_host(const double *terms_host, int64_t n, double *out_host) {
    MSS_REQUIRE(terms_host && out_host && n >= 1, "mss_pairwise_sum_host: bad arguments");
    PwPlan p;
    MSS_REQUIRE(pw_plan(n, p), "mss_pairwise_sum_host: plan bound exceeded");
    PwTree tree{p.size.data(), p.leaves.data(), (int)p.size.size(), p.n, p.n_leaves};
    std::vector<double> leaf((size_t)p.n_leaves);
    for (long long l = 0; l < p.n_leaves; l++) {
        long long s, m;
        tree.leaf_bounds(l, s, m);
        const double *a = terms_host + s;
        volatile double res;
        if (m < 8) {
            res = -0.0;
            for (long long i = 0; i < m; i++) res = res + a[i];
        } else {
            volatile double r[8];
            for (int j = 0; j < 8; j++) r[j] = a[j];
            long long i;
            for (i = 8; i < m - (m % 8); i += 8)
                for (int j = 0; j < 8; j++) r[j] = r[j] + a[i + j];
            volatile double q0 = r[0] + r[1], q1 = r[2] + r[3], q2 = r[4] + r[5], q3 = r[6] + r[7];
            volatile double h0 = q0 + q1, h1 = q2 + q3;
            res = h0 + h1;
            for (; i < m; i++) res = res + a[i];
        }
        leaf[(size_t)l] = res;
    }
    std::vector<double> node(p.front.size());
    for (size_t i = 0; i < p.front.size(); i++) node[i] = pw_subtree_combine(p.front[i].m, leaf.data(), p.front[i].first_leaf);
    size_t nx = 0;
    *out_host = pw_combine_top(node.data(), nx, n, p.F);
    MSS_REQUIRE(nx == node.size(), "mss_pairwise_sum_host: frontier mismatch");
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
