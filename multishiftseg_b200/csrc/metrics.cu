// (a9-a13) exact tie-aware AUROC / AP / FPR@95TPR from the two sorted key streams of an evaluation
// (negatives = in-distribution keys, positives = OOD keys; ascending key == descending float32 score).
//
// Integer stage (bit-exact by construction):
//   thresholds = distinct keys of the merged streams (== np.where(np.diff(y_score)) + last index,
//   sklearn _ranking.py:916-920, metric.py:110-111);  tps[k] = #positives with key <= threshold k
//   (cumsum(y)[idx], _ranking.py:1034 / metric.py:114), fps[k] = #negatives with key <= threshold k
//   (= 1 + idx - tps).  One merge-path pass over both streams, one decoupled look-back for the compaction
//   offset: every key is read once, every threshold written once (round 1: a count pass, a single-CTA scan
//   and a write pass over (key, label) pairs).
// float64 tail (must reproduce numpy/sklearn rounding exactly, so: explicit __d*_rn intrinsics, no
// FMA contraction, and numpy's pairwise summation tree replayed leaf by leaf):
//   AUROC  roc_curve(drop_intermediate=True) + auc/trapezoid   _ranking.py:1331-1378, :53-116
//   AP     precision_recall_curve + step integral              _ranking.py:1160-1208, :243-260
//   FPR95  fpr_and_fdr_at_recall                               metric.py:116-127
// This translation unit is compiled with -fmad=false.
#include <map>
#include <vector>

#include "sort_plan.cuh"

namespace mss {

constexpr int CT_THREADS = 256;

// block-wide exclusive scan of one count per thread; CTA total in `tot`
// BAR = 0: __syncthreads (every thread of the CTA takes part); BAR > 0: named barrier over the first THREADS threads (the
// CTA also has a walker warp that must not be waited for)
template <int THREADS = CT_THREADS, int BAR = 0>
__device__ __forceinline__ void scan_barrier() {
    if (BAR == 0) __syncthreads(); else named_barrier(BAR, THREADS);
}
template <int THREADS = CT_THREADS, int BAR = 0>
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned &tot) {
    __shared__ unsigned s_w[THREADS / 32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    scan_barrier<THREADS, BAR>();
    unsigned base = 0, all = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) {
        const unsigned x = s_w[w];
        if (w < (int)warp) base += x;
        all += x;
    }
    tot = all;
    scan_barrier<THREADS, BAR>();
    return base + inc - v;
}

// ---- sorted streams -> (tps, fps) -----------------------------------------------------------------
// Merge path: among the first `diag` keys of the merged order (ties: negatives first -- the order inside a tie
// does not matter, a threshold is emitted at the END of a run of equal keys), how many come from A?
__device__ __forceinline__ long long merge_path(const uint32_t *__restrict__ A, long long nA,
                                                const uint32_t *__restrict__ B, long long nB, long long diag) {
    long long lo = diag > nB ? diag - nB : 0, hi = diag < nA ? diag : nA;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(A + mid) <= __ldg(B + (diag - 1 - mid))) lo = mid + 1; else hi = mid;
    }
    return lo;
}

constexpr int MC_THREADS = 256;
constexpr int MC_IPT = 16;
constexpr int MC_TILE = MC_THREADS * MC_IPT;   // 4096 merged keys per CTA: half as many look-backs as 2048

// one thread per tile boundary: a_start[t] = merge path at diagonal min(t * MC_TILE, nA + nB), t = 0 .. tiles_upper
__global__ void __launch_bounds__(256)
merge_partition_kernel(const SortPlan *__restrict__ plan, long long *__restrict__ a_start, long long tiles_upper) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > tiles_upper) return;
    const long long nA = plan->seg[0].n, nB = plan->seg[1].n;
    const long long diag = min(t * MC_TILE, nA + nB);
    a_start[t] = merge_path(plan->seg[0].x, nA, plan->seg[1].x, nB, diag);
}

// totals[0] = T (distinct keys), written by the last tile.
// A threshold is staged as two 16-bit tile-local numbers (negatives consumed, merged position) and widened to the int64
// (tps, fps) pair only in the coalesced write-out.
__global__ void __launch_bounds__(MC_THREADS + 32)
merge_counts_kernel(const SortPlan *__restrict__ plan, const long long *__restrict__ a_start, long long pos_before,
                    long long neg_before, long long *__restrict__ tps, long long *__restrict__ fps,
                    unsigned long long *status, unsigned *counter, unsigned long long *__restrict__ totals) {
    __shared__ uint32_t s_k[MC_TILE];
    __shared__ unsigned short s_a[MC_TILE], s_p[MC_TILE];
    __shared__ unsigned s_tile, s_tot;
    __shared__ unsigned long long s_excl;
    const unsigned tid = threadIdx.x;
    if (tid == 0) s_tile = atomicAdd(counter, 1u);       // tiles in start order: a walk never waits on a tile not yet running
    __syncthreads();
    const unsigned tile = s_tile;
    const long long nA = plan->seg[0].n, nB = plan->seg[1].n, M = nA + nB;
    const long long tiles = (M + MC_TILE - 1) / MC_TILE;
    if ((long long)tile >= tiles) return;
    if (tid >= MC_THREADS) {
        // ---- walker warp: the exclusive prefix, while the workers merge
        const unsigned long long ex = tile_walk(status, tile);
        __syncthreads();                                                  // (A) the tile's count is known
        if (tid == MC_THREADS) {
            tile_publish_inclusive(status, tile, ex + s_tot);
            s_excl = ex;
            if ((long long)tile == tiles - 1) totals[0] = ex + s_tot;
        }
        __syncthreads();                                                  // (B)
        return;
    }
    const uint32_t *__restrict__ A = plan->seg[0].x, *__restrict__ B = plan->seg[1].x;
    const long long d0 = (long long)tile * MC_TILE, d1 = min(M, d0 + MC_TILE);
    const long long a0 = a_start[tile], a1 = a_start[tile + 1], b0 = d0 - a0, b1 = d1 - a1;
    const int na = (int)(a1 - a0), nb = (int)(b1 - b0), cnt = na + nb;
    {
        const uint32_t *pa = A + a0, *pb = B + b0 - na;
        for (int i = tid; i < cnt; i += MC_THREADS) s_k[i] = (i < na) ? __ldg(pa + i) : __ldg(pb + i);
    }
    // the key that follows this tile in the merged order (decides whether the tile's last key ends a run)
    const bool has_next = d1 < M;
    uint32_t nxt = 0;
    if (has_next) {
        const uint32_t ka = (a1 < nA) ? __ldg(A + a1) : 0xFFFFFFFFu, kb = (b1 < nB) ? __ldg(B + b1) : 0xFFFFFFFFu;
        nxt = (a1 < nA && b1 < nB) ? min(ka, kb) : (a1 < nA ? ka : kb);
    }
    named_barrier(1, MC_THREADS);

    // this thread's MC_IPT keys of the merged order start at diagonal `diag` of the tile
    const int diag = min((int)tid * MC_IPT, cnt);
    int lo = max(0, diag - nb), hi = min(diag, na);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_k[mid] <= s_k[na + diag - 1 - mid]) lo = mid + 1; else hi = mid;
    }
    int ai = lo, bi = diag - lo;
    const int ai0 = ai;
    unsigned took_a = 0, ends = 0;
    {
        // heads of the two runs in registers; only the consumed one is re-read
        uint32_t ka = (ai < na) ? s_k[ai] : 0u, kb = (bi < nb) ? s_k[na + bi] : 0u;
        bool ta = (bi >= nb) || (ai < na && ka <= kb);
        uint32_t key = ta ? ka : kb;
#pragma unroll
        for (int j = 0; j < MC_IPT; j++) {
            const int p = diag + j;
            if (p < cnt) {
                if (ta) { ai++; took_a |= 1u << j; if (ai < na) ka = s_k[ai]; }
                else { bi++; if (bi < nb) kb = s_k[na + bi]; }
                bool end;
                if (p + 1 < cnt) {
                    ta = (bi >= nb) || (ai < na && ka <= kb);
                    const uint32_t nk = ta ? ka : kb;
                    end = nk != key;
                    key = nk;
                } else {
                    end = !has_next || nxt != key;
                }
                if (end) ends |= 1u << j;
            }
        }
    }
    unsigned tot;
    unsigned slot = block_exclusive_scan<MC_THREADS, 1>(__popc(ends), tot);
    if (tid == 0) {
        tile_publish_aggregate(status, tile, tot);       // the successors' walkers can use it long before our own walk ends
        s_tot = tot;
    }
#pragma unroll
    for (int j = 0; j < MC_IPT; j++) {
        if ((ends >> j) & 1u) {
            s_a[slot] = (unsigned short)(ai0 + __popc(took_a & ((2u << j) - 1u)));   // negatives consumed up to and including item j
            s_p[slot] = (unsigned short)(diag + j + 1);                              // merged keys consumed (<= 4096)
            slot++;
        }
    }
    __syncthreads();                                                      // (A)
    __syncthreads();                                                      // (B) the walker has stored the prefix
    long long *__restrict__ tp = tps + s_excl, *__restrict__ fp = fps + s_excl;
    const long long tb = pos_before + b0, fb = neg_before + a0;
    for (unsigned q = tid; q < tot; q += MC_THREADS) {
        const int aj = s_a[q], pj = s_p[q];
        tp[q] = tb + (pj - aj);
        fp[q] = fb + aj;
    }
}

// ---- ROC: drop collinear points (roc_curve drop_intermediate, _ranking.py:1338-1350) ----------------
// ---- FPR@95: argmin_k |tps[k]/P - 0.95| over k <= first k with tps[k]==P, ties -> largest k -----------
//      (fpr_and_fdr_at_recall, metric.py:116-127; fused into the compaction pass below: same data window)
struct Best {
    double d;
    long long k;
};
__device__ __forceinline__ Best better(Best x, Best y) {
    if (y.d < x.d || (y.d == x.d && y.k > x.k)) return y;
    return x;
}
template <int THREADS = CT_THREADS, int BAR = 0>
__device__ __forceinline__ Best block_best(Best b) {      // valid in thread 0
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        Best o{__shfl_xor_sync(0xffffffffu, b.d, s), __shfl_xor_sync(0xffffffffu, b.k, s)};
        b = better(b, o);
    }
    __shared__ Best sb[THREADS / 32];
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = b;
    scan_barrier<THREADS, BAR>();
    if (threadIdx.x == 0)
        for (int w = 1; w < THREADS / 32; w++) b = better(b, sb[w]);
    return b;
}

// T comes from the device (totals of the counting pass) when d_T is given, else from the host.
__device__ __forceinline__ long long load_T(const unsigned long long *d_T, long long T_host) {
    return d_T ? (long long)*d_T : T_host;
}

// Single pass over DRAM, fat tiles: a CTA owns up to 16384 consecutive thresholds as `sub` <= RC_SUB sub-tiles of
// 256 x 4 thresholds (the host picks `sub` so that small inputs still fill the GPU).
//   phase A  per sub-tile: keep flags (second differences over a window i0-1 .. i0+4, the neighbours' values by
//            shuffle) and the FPR95 candidate -- a pure streaming loop, no barrier inside, so the loads of several
//            sub-tiles are in flight; the 4-bit keep masks stay in one 64-bit register;
//   offsets  ONE multi-scan for all sub-tiles at the end: nibble popcounts by SWAR, byte-packed warp scans (16 counters in
//            two 64-bit words), a 128-entry (sub-tile, warp) table scanned by one warp -- two barriers per tile;
//   walk     a dedicated warp computes the exclusive prefix over the previous tiles meanwhile (one look-back per tile);
//   phase B  the tile is read again -- from L2, the CTAs of one wave hold < 80 MB -- and the kept points are written at
//            their final offsets.
// History (round 2, T = 67 M): 256 x 8 thresholds per CTA with warp 0 walking after the scan 1.20 ms (9 polls of ~0.5 us
// per tile, 55 % of the stall samples on the barrier behind the walk); 16384-threshold tiles with a scan per sub-tile
// 0.88 ms (latency-bound: a load round trip and two barriers per sub-tile); this form: see profiles/.
// totals_out[0] = number of kept points (written by the last tile); tile_best[tile] = FPR95 candidate of the tile.
constexpr int RC_THREADS = 256;
constexpr int RC_IPT = 4;
constexpr int RC_SUB = 16;                                // layout bound of the (sub-tile, warp) tables: 16 x 8 = 32 lanes x 4
constexpr int RC_SUB_USED = 8;                            // tiles of 8192 thresholds: with 16 sub-tiles the re-read of phase B came
                                                          // from DRAM (ncu: 2.0 GB read for 1.05 GB of input)
constexpr int RC_SUBTILE = RC_THREADS * RC_IPT;           // 1024 thresholds
constexpr int CT_TILE_MAX = RC_SUBTILE * RC_SUB;          // 16384 thresholds per ROC / FPR95 tile at most

__device__ __forceinline__ void roc_load4(const long long *__restrict__ tps, const long long *__restrict__ fps, long long i0,
                                          long long T, long long (&t)[RC_IPT + 2], long long (&f)[RC_IPT + 2]) {
    if (i0 + RC_IPT <= T && ((((uintptr_t)tps) | ((uintptr_t)fps)) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < RC_IPT; j += 2) {
            const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(tps + i0 + j));
            const longlong2 c = __ldg(reinterpret_cast<const longlong2 *>(fps + i0 + j));
            t[j + 1] = a.x; t[j + 2] = a.y;
            f[j + 1] = c.x; f[j + 2] = c.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < RC_IPT; j++) {
            const bool in = i0 + j < T;
            t[j + 1] = in ? __ldg(tps + i0 + j) : 0;
            f[j + 1] = in ? __ldg(fps + i0 + j) : 0;
        }
    }
}

__global__ void __launch_bounds__(RC_THREADS + 32)
roc_compact_kernel(const long long *__restrict__ tps, const long long *__restrict__ fps, const unsigned long long *d_T,
                   long long T_host, int sub, double recall_level, unsigned long long *status, unsigned *counter,
                   Best *__restrict__ tile_best, long long *__restrict__ tps_k, long long *__restrict__ fps_k,
                   unsigned long long *__restrict__ totals_out) {
    __shared__ unsigned s_tile, s_tot;
    __shared__ unsigned long long s_excl;
    __shared__ unsigned char s_wt[RC_SUB][RC_THREADS / 32];      // kept points of (sub-tile, warp)
    __shared__ unsigned short s_base[RC_SUB][RC_THREADS / 32];   // their exclusive prefix in (sub-tile, warp) order
    const unsigned tid = threadIdx.x;
    if (tid == 0) s_tile = atomicAdd(counter, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const long long T = load_T(d_T, T_host);
    const long long tile_thr = (long long)sub * RC_SUBTILE;
    const long long tiles = (T + tile_thr - 1) / tile_thr;
    if ((long long)tile >= tiles) return;
    if (tid >= RC_THREADS) {
        const unsigned long long ex = tile_walk(status, tile);
        __syncthreads();                                                  // (A) the tile's count is known
        if (tid == RC_THREADS) {
            tile_publish_inclusive(status, tile, ex + s_tot);
            s_excl = ex;
            if ((long long)tile == tiles - 1) totals_out[0] = ex + s_tot;
        }
        __syncthreads();                                                  // (B)
        return;
    }
    const unsigned lane = tid & 31, warp = tid >> 5;
    const long long tile0 = (long long)tile * tile_thr;
    const double P = (double)__ldg(tps + T - 1);
    unsigned long long masks = 0;                  // 4 keep bits per sub-tile
    Best best{INFINITY, -1};
#pragma unroll 2
    for (int sb = 0; sb < sub; sb++) {
        const long long i0 = tile0 + (long long)sb * RC_SUBTILE + (long long)tid * RC_IPT;
        long long t[RC_IPT + 2], f[RC_IPT + 2];                  // window index w <-> threshold i0 - 1 + w
        roc_load4(tps, fps, i0, T, t, f);
        // the neighbours' values travel by shuffle; only the warp's edge lanes read them from memory
        {
            const long long tl = __shfl_up_sync(0xffffffffu, t[RC_IPT], 1), fl = __shfl_up_sync(0xffffffffu, f[RC_IPT], 1);
            const long long tr = __shfl_down_sync(0xffffffffu, t[1], 1), fr = __shfl_down_sync(0xffffffffu, f[1], 1);
            if (lane == 0) {
                t[0] = (i0 > 0 && i0 - 1 < T) ? __ldg(tps + i0 - 1) : 0;
                f[0] = (i0 > 0 && i0 - 1 < T) ? __ldg(fps + i0 - 1) : 0;
            } else { t[0] = tl; f[0] = fl; }
            if (lane == 31) {
                t[RC_IPT + 1] = (i0 + RC_IPT < T) ? __ldg(tps + i0 + RC_IPT) : 0;
                f[RC_IPT + 1] = (i0 + RC_IPT < T) ? __ldg(fps + i0 + RC_IPT) : 0;
            } else { t[RC_IPT + 1] = tr; f[RC_IPT + 1] = fr; }
        }
        unsigned keep = 0;
#pragma unroll
        for (int j = 0; j < RC_IPT; j++) {
            const long long k = i0 + j;
            if (k < T) {
                bool kp = true;
                if (T > 2 && k != 0 && k != T - 1) {
                    const long long d2f = f[j + 2] - 2 * f[j + 1] + f[j];
                    const long long d2t = t[j + 2] - 2 * t[j + 1] + t[j];
                    kp = d2f != 0 || d2t != 0;
                }
                if (kp) keep |= 1u << j;
                if (k == 0 || (double)t[j] != P) {     // FPR95: k <= searchsorted(tps, tps[-1])
                    const double d = fabs(__dsub_rn(__ddiv_rn((double)t[j + 1], P), recall_level));
                    best = better(best, Best{d, k});
                }
            }
        }
        masks |= (unsigned long long)keep << (4 * sb);
    }
    // ---- offsets of all sub-tiles at once.  Nibble popcounts (SWAR): nibble sb of `pc` = kept points of sub-tile sb.
    unsigned long long pc = masks - ((masks >> 1) & 0x5555555555555555ull);
    pc = (pc & 0x3333333333333333ull) + ((pc >> 2) & 0x3333333333333333ull);
    // bytes: word e holds the even sub-tiles (byte q <-> sub-tile 2q), word o the odd ones; inclusive warp scans
    // (a warp keeps <= 128 points of a sub-tile: fits a byte)
    unsigned long long e = pc & 0x0F0F0F0F0F0F0F0Full, o = (pc >> 4) & 0x0F0F0F0F0F0F0F0Full;
    const unsigned long long own_e = e, own_o = o;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long te = __shfl_up_sync(0xffffffffu, e, d), to = __shfl_up_sync(0xffffffffu, o, d);
        if (lane >= (unsigned)d) { e += te; o += to; }
    }
    if (lane == 31) {
#pragma unroll
        for (int q = 0; q < RC_SUB / 2; q++) {
            s_wt[2 * q][warp] = (unsigned char)(e >> (8 * q));
            s_wt[2 * q + 1][warp] = (unsigned char)(o >> (8 * q));
        }
    }
    best = block_best<RC_THREADS, 1>(best);                              // (contains a worker barrier: s_wt is complete after it)
    if (tid == 0) tile_best[tile] = best;
    if (warp == 0) {
        // 128 (sub-tile, warp) counts in sub-tile-major order, 4 per lane
        const unsigned char *w = &s_wt[0][0] + lane * 4;
        const unsigned c0 = w[0], c1 = w[1], c2 = w[2], c3 = w[3];
        unsigned inc = c0 + c1 + c2 + c3;
        const unsigned sum = inc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned tt = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (unsigned)d) inc += tt;
        }
        unsigned short *bo = &s_base[0][0] + lane * 4;
        const unsigned x0 = inc - sum;
        bo[0] = (unsigned short)x0; bo[1] = (unsigned short)(x0 + c0); bo[2] = (unsigned short)(x0 + c0 + c1);
        bo[3] = (unsigned short)(x0 + c0 + c1 + c2);
        if (lane == 31) {
            tile_publish_aggregate(status, tile, inc);
            s_tot = inc;
        }
    }
    __syncthreads();                                                      // (A)
    __syncthreads();                                                      // (B) the walker has stored the prefix
    const unsigned long long base = s_excl;
    e -= own_e;                                                           // exclusive inside the warp
    o -= own_o;
#pragma unroll 2
    for (int sb = 0; sb < sub; sb++) {
        const unsigned keep = (unsigned)(masks >> (4 * sb)) & 15u;
        if (keep) {
            const long long i0 = tile0 + (long long)sb * RC_SUBTILE + (long long)tid * RC_IPT;
            long long t[RC_IPT + 2], f[RC_IPT + 2];
            roc_load4(tps, fps, i0, T, t, f);                      // second read of the tile: L2
            const unsigned in_warp = (unsigned)(((sb & 1) ? o : e) >> (8 * (sb >> 1))) & 255u;
            unsigned long long op = base + s_base[sb][warp] + in_warp;
#pragma unroll
            for (int j = 0; j < RC_IPT; j++) {
                if ((keep >> j) & 1u) {
                    tps_k[op] = t[j + 1];
                    fps_k[op] = f[j + 1];
                    op++;
                }
            }
        }
    }
}

// reduce the tile candidates to gridDim.x candidates (grid-stride)
__global__ void __launch_bounds__(CT_THREADS)
fpr_reduce_kernel(const Best *__restrict__ in, const unsigned long long *d_T, long long T_host, long long tile_thr,
                  Best *__restrict__ out) {
    const long long T = load_T(d_T, T_host), n = (T + tile_thr - 1) / tile_thr;
    Best b{INFINITY, -1};
    for (long long i = (long long)blockIdx.x * CT_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * CT_THREADS)
        b = better(b, in[i]);
    b = block_best(b);
    if (threadIdx.x == 0) out[blockIdx.x] = b;
}

// n_partial > 0: `cand` holds n_partial reduced candidates; else it holds one candidate per tile
__global__ void __launch_bounds__(CT_THREADS)
fpr_final_kernel(const Best *__restrict__ cand, int n_partial, const long long *__restrict__ fps,
                 const unsigned long long *d_T, long long T_host, long long tile_thr, double *__restrict__ out) {
    const long long T = load_T(d_T, T_host);
    const long long n = n_partial > 0 ? n_partial : (T + tile_thr - 1) / tile_thr;
    Best b{INFINITY, -1};
    for (long long i = threadIdx.x; i < n; i += CT_THREADS) b = better(b, cand[i]);
    b = block_best(b);
    if (threadIdx.x == 0) {
        const double N = T > 0 ? (double)fps[T - 1] : 0.0;
        out[0] = (b.k >= 0) ? __ddiv_rn((double)fps[b.k], N) : nan("");   // k < 0 only when P == 0 (caller reports it)
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

// one thread per leaf: start offsets of all leaves (leaf_start[n_leaves] = n); blockIdx.y selects the tree (AP / ROC)
struct PwTree2 {
    PwTree tree[2];
    long long *leaf_start[2];
};
__global__ void __launch_bounds__(256)
leaf_bounds_kernel(PwTree2 tt) {
    const PwTree &tree = tt.tree[blockIdx.y];
    const long long leaf = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf > tree.n_leaves) return;
    long long s = tree.n, m = 0;
    if (leaf < tree.n_leaves) tree.leaf_bounds(leaf, s, m);
    tt.leaf_start[blockIdx.y][leaf] = s;
}

// 8 lanes per leaf = numpy's 8 interleaved accumulators
template <typename Term>
__device__ __forceinline__ void leaf_sum_body(const Term &term, const long long *__restrict__ leaf_start, long long n_leaves,
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

// Leaf sums of both integrals.  A CTA owns LS_LEAVES consecutive leaves (<= 128 terms each), i.e. ONE contiguous range of
// thresholds: it is staged into shared memory as float64 by coalesced loads issued all at once (the first form read
// tps[k] / fps[k] per term inside the summation loop and ncu showed 70 % of the stall samples waiting on those loads,
// issue 27 %, DRAM 21 %), then 8 lanes per leaf -- numpy's 8 interleaved accumulators -- sum from shared memory.
// Neighbouring terms share float64 quotients (AP term j needs rec[k] and rec[k-1], the ROC term needs point j and point
// j-1): a lane computes only ITS OWN quotients and takes the neighbour's through a shuffle -- 2 instead of 3 resp. 4
// float64 divides per term; values and operation order per term are unchanged, hence so is every bit of the sums.
//   AP : lane sub of block i owns k = K0 - i - sub; rec[k-1] is lane sub+1's value, and for lane 7 lane 0's value of the
//        NEXT block, which is computed one block ahead anyway.
//   ROC: lane sub owns point j; point j-1 is lane sub-1's value, and for lane 0 lane 7's value of the previous block.
constexpr int LS_THREADS = 256;
constexpr int LS_LEAVES = LS_THREADS / 8;                 // 32 leaves per CTA
constexpr int LS_MAX_TERMS = LS_LEAVES * 128;             // 4096
constexpr int LS_STAGE = LS_MAX_TERMS + 8;                // + the neighbour point
constexpr int LS_SMEM = 2 * LS_STAGE * 8;

// blockIdx.y = 0: AP leaves, 1: ROC leaves (one launch for both sums)
struct LeafJob {
    ApTerm ap;
    RocTerm roc;
    const long long *leaf_start[2];
    long long n_leaves[2];
    double *leaf_sum[2];
};

__global__ void __launch_bounds__(LS_THREADS)
leaf_sum_kernel(LeafJob job) {
    extern __shared__ __align__(16) double s_stage[];
    double *s_t = s_stage, *s_f = s_stage + LS_STAGE;
    const int y = blockIdx.y;
    const long long n_leaves = job.n_leaves[y];
    const long long first = (long long)blockIdx.x * LS_LEAVES;
    if (first >= n_leaves) return;
    const long long *__restrict__ leaf_start = job.leaf_start[y];
    const long long last = min(first + LS_LEAVES, n_leaves);
    const long long S0 = leaf_start[first], S1 = leaf_start[last];          // terms [S0, S1) of this CTA
    const unsigned tid = threadIdx.x, sub = tid & 7;
    const unsigned gmask = 0xffu << (tid & 24);     // the 8 lanes of this leaf: leaves of one warp differ in length
    const long long leaf = first + (tid >> 3);
    const bool live = leaf < last;
    const long long s = live ? leaf_start[leaf] : S0;
    const long long m = live ? leaf_start[leaf + 1] - s : 0;
    double res = -0.0;
    if (y == 0) {
        // ---- AP: term j <-> threshold k = T-1-j; staged thresholds k in [klo, khi], khi = T-1-S0, klo = T-1-S1 (>= -1)
        const ApTerm &term = job.ap;
        const long long T = term.T, klo = T - 1 - S1, khi = T - 1 - S0;
        for (long long q = tid; q <= khi - klo; q += LS_THREADS) {
            const long long k = klo + q;
            s_t[q] = k >= 0 ? (double)__ldg(term.tps + k) : 0.0;
            s_f[q] = k >= 0 ? (double)__ldg(term.fps + k) : 0.0;
        }
        const double P = (double)__ldg(term.tps + T - 1);
        __syncthreads();
        // rec[-1] := 0 (the appended (1, 0) point of precision_recall_curve)
        // (k = -1 is staged as 0, and 0 / P = +0; k < klo only occurs for the look-ahead of lanes that do not use it)
        auto rec = [&](long long k) { return k >= klo ? __ddiv_rn(s_t[k - klo], P) : 0.0; };
        auto termv = [&](long long j) {
            const long long k = T - 1 - j;
            const double t = s_t[k - klo], f = s_f[k - klo];
            return __dmul_rn(__dsub_rn(rec(k - 1), rec(k)), __ddiv_rn(t, __dadd_rn(t, f)));
        };
        if (m < 8) {
            if (sub == 0)
                for (long long i = 0; i < m; i++) res = __dadd_rn(res, termv(s + i));
        } else {
            const long long body = m - (m & 7);
            const long long K0 = T - 1 - s - sub;                        // this lane's k in block 0
            double r_cur = rec(K0), r = 0.0;
            for (long long i = 0; i < body; i += 8) {
                const double r_next = rec(K0 - i - 8);                   // next block's own quotient
                double r_prev = __shfl_down_sync(gmask, r_cur, 1, 8);
                const double r_wrap = __shfl_sync(gmask, r_next, 0, 8);
                if (sub == 7) r_prev = r_wrap;
                const long long k = K0 - i;
                const double t = s_t[k - klo], f = s_f[k - klo];
                const double v = __dmul_rn(__dsub_rn(r_prev, r_cur), __ddiv_rn(t, __dadd_rn(t, f)));
                r = (i == 0) ? v : __dadd_rn(r, v);
                r_cur = r_next;
            }
            // ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)) -- fp addition is commutative, so the butterfly is exact
            r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 1));
            r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 2));
            r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 4));
            res = r;
            if (sub == 0)
                for (long long i = body; i < m; i++) res = __dadd_rn(res, termv(s + i));
        }
    } else {
        // ---- ROC: term j joins points j-1 and j of the kept curve (point -1 = the origin); staged points [S0-1, S1-1]
        const RocTerm &term = job.roc;
        const long long jlo = S0 - 1;
        for (long long q = tid; q <= S1 - 1 - jlo; q += LS_THREADS) {
            const long long j = jlo + q;
            s_t[q] = j >= 0 ? (double)__ldg(term.tps_k + j) : 0.0;
            s_f[q] = j >= 0 ? (double)__ldg(term.fps_k + j) : 0.0;
        }
        const double P = (double)__ldg(term.p_P), N = (double)__ldg(term.p_N);
        __syncthreads();
        auto termv = [&](long long j) {
            const double f1 = __ddiv_rn(s_f[j - jlo], N), t1 = __ddiv_rn(s_t[j - jlo], P);
            const double f0 = __ddiv_rn(s_f[j - 1 - jlo], N), t0 = __ddiv_rn(s_t[j - 1 - jlo], P);
            return __ddiv_rn(__dmul_rn(__dsub_rn(f1, f0), __dadd_rn(t1, t0)), 2.0);
        };
        if (m < 8) {
            if (sub == 0)
                for (long long i = 0; i < m; i++) res = __dadd_rn(res, termv(s + i));
        } else {
            const long long body = m - (m & 7);
            // point s - 1: needed by lane 0 of the first block only
            double f_carry = __ddiv_rn(s_f[s - 1 - jlo], N), t_carry = __ddiv_rn(s_t[s - 1 - jlo], P);
            double r = 0.0;
            for (long long i = 0; i < body; i += 8) {
                const long long j = s + i + sub;
                const double f1 = __ddiv_rn(s_f[j - jlo], N), t1 = __ddiv_rn(s_t[j - jlo], P);
                double f0 = __shfl_up_sync(gmask, f1, 1, 8), t0 = __shfl_up_sync(gmask, t1, 1, 8);
                if (sub == 0) { f0 = f_carry; t0 = t_carry; }
                const double v = __ddiv_rn(__dmul_rn(__dsub_rn(f1, f0), __dadd_rn(t1, t0)), 2.0);
                r = (i == 0) ? v : __dadd_rn(r, v);
                f_carry = __shfl_sync(gmask, f1, 7, 8);                  // lane 7's point is lane 0's predecessor next block
                t_carry = __shfl_sync(gmask, t1, 7, 8);
            }
            r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 1));
            r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 2));
            r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 4));
            res = r;
            if (sub == 0)
                for (long long i = body; i < m; i++) res = __dadd_rn(res, termv(s + i));
        }
    }
    if (live && sub == 0) job.leaf_sum[y][leaf] = res;
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

struct CombineJob {
    const PwFrontNode *nodes[2];
    int n_nodes[2];
    const double *leaf_sum[2];
    double *node_sum[2];
};
__global__ void __launch_bounds__(128)
subtree_combine_kernel(CombineJob job) {
    const int y = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= job.n_nodes[y]) return;
    job.node_sum[y][i] = pw_subtree_combine(job.nodes[y][i].m, job.leaf_sum[y], job.nodes[y][i].first_leaf);
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

static size_t ct_tiles(int64_t n) { return (size_t)((n + RC_SUBTILE - 1) / RC_SUBTILE); }    // sizing bound: one sub-tile per tile
static size_t mc_tiles(int64_t n) { return (size_t)((n + MC_TILE - 1) / MC_TILE); }

// ---- counting stage ----------------------------------------------------------------------------------
struct CountsWs {
    SortPlan *plan;                // used when the caller has no plan of its own (mss_counts_from_sorted)
    long long *a_start;            // [tiles_upper + 1] merge-path splits at the tile boundaries
    unsigned *counter;             // tile ticket           } zeroed together
    unsigned long long *totals;    // [0] = T               }
    unsigned long long *status;    // [tiles_upper]         }
    char *zero_base;
    size_t zero_bytes, tiles_upper;
};
static bool carve_counts(void *ws, size_t bytes, int64_t n_upper, CountsWs &o) {
    Carver c(ws, bytes);
    o.tiles_upper = mc_tiles(n_upper);
    o.plan = c.take<SortPlan>(1);
    o.a_start = c.take<long long>(o.tiles_upper + 1);
    o.counter = c.take<unsigned>(64);
    o.zero_base = (char *)o.counter;
    o.totals = c.take<unsigned long long>(4);
    o.status = c.take<unsigned long long>(o.tiles_upper + 1);
    o.zero_bytes = (size_t)((char *)(o.status + o.tiles_upper + 1) - o.zero_base);
    return c.ok();
}
static size_t counts_ws_bytes(int64_t n) {
    if (n < 0) n = 0;
    return 256 + align_up((mc_tiles(n) + 1) * 8, 256) + 256 + 256 + align_up((mc_tiles(n) + 1) * 8, 256) + 1024;
}

// merge the plan's two sorted streams into per-threshold cumulative counts; T lands in w.totals[0] (device)
static int counts_enqueue(const SortPlan *plan, int64_t n_upper, int64_t pos_before, int64_t neg_before, int64_t *tps,
                          int64_t *fps, const CountsWs &w, cudaStream_t st) {
    MSS_REQUIRE(w.tiles_upper < (1ull << 31), "counts: n too large");
    MSS_CHECK_CUDA(cudaMemsetAsync(w.zero_base, 0, w.zero_bytes, st));
    if (w.tiles_upper == 0) return MSS_OK;
    merge_partition_kernel<<<(unsigned)((w.tiles_upper + 1 + 255) / 256), 256, 0, st>>>(plan, w.a_start, (long long)w.tiles_upper);
    MSS_CHECK_LAUNCH();
    merge_counts_kernel<<<(unsigned)w.tiles_upper, MC_THREADS + 32, 0, st>>>(plan, w.a_start, pos_before, neg_before, (long long *)tps,
                                                                       (long long *)fps, w.status, w.counter, w.totals);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

// ---- float64 tail --------------------------------------------------------------------------------------
struct TailWs {
    long long *tps_k, *fps_k;       // [T_upper] points kept by roc_curve(drop_intermediate=True)
    Best *tile_best, *partial;      // FPR95 candidates per tile / reduced
    double *sum_ap, *sum_roc;       // leaf sums
    long long *leaf_start[2];
    char *blob;                     // tree tables + frontier nodes of both trees (one H2D copy)
    double *results;                // [0] = FPR95, then the frontier sums of both trees (one D2H copy)
    unsigned *counter;              // } zeroed together
    unsigned long long *totals;     // } [0] = T_roc
    unsigned long long *status;     // } [tiles_upper + 1]
    char *zero_base;
    size_t zero_bytes, tiles_upper, max_leaves;
};
static size_t pw_blob_bytes() { return 2 * (2 * (size_t)PW_MAX_SIZES * 8 + (size_t)PW_MAX_FRONT * sizeof(PwFrontNode)); }
static bool carve_tail(void *ws, size_t bytes, int64_t T_upper, TailWs &o) {
    Carver c(ws, bytes);
    o.tiles_upper = ct_tiles(T_upper);
    o.max_leaves = (size_t)T_upper / 64 + 2;
    o.tps_k = c.take<long long>((size_t)T_upper);
    o.fps_k = c.take<long long>((size_t)T_upper);
    o.tile_best = c.take<Best>(o.tiles_upper + 1);
    o.partial = c.take<Best>(1024);
    o.sum_ap = c.take<double>(o.max_leaves);
    o.sum_roc = c.take<double>(o.max_leaves);
    o.leaf_start[0] = c.take<long long>(o.max_leaves + 1);
    o.leaf_start[1] = c.take<long long>(o.max_leaves + 1);
    o.blob = c.take<char>(pw_blob_bytes());
    o.results = c.take<double>(2 * PW_MAX_FRONT + 8);
    o.counter = c.take<unsigned>(64);
    o.zero_base = (char *)o.counter;
    o.totals = c.take<unsigned long long>(4);
    o.status = c.take<unsigned long long>(o.tiles_upper + 1);
    o.zero_bytes = (size_t)((char *)(o.status + o.tiles_upper + 1) - o.zero_base);
    return c.ok();
}
static size_t tail_ws_bytes(int64_t T) {
    if (T < 0) T = 0;
    const size_t leaves = (size_t)T / 64 + 2, tiles = ct_tiles(T);
    return 2 * align_up((size_t)T * 8, 256) + align_up((tiles + 1) * sizeof(Best), 256) + align_up(1024 * sizeof(Best), 256) +
           2 * align_up(leaves * 8, 256) + 2 * align_up((leaves + 1) * 8, 256) + align_up(pw_blob_bytes(), 256) +
           align_up((2 * PW_MAX_FRONT + 8) * 8, 256) + 256 + 256 + align_up((tiles + 1) * 8, 256) + 2048;
}

// stage 1 (no host data needed): ROC compaction + FPR95.  T is read from the device when d_T is given.
// T_roc lands in w.totals[0], FPR95 in w.results[0].
static int tail1_enqueue(const int64_t *tps_, const int64_t *fps_, const unsigned long long *d_T, int64_t T_host,
                         double recall_level, const TailWs &w, cudaStream_t st) {
    const long long *tps = (const long long *)tps_, *fps = (const long long *)fps_;
    MSS_REQUIRE(w.tiles_upper < (1ull << 31), "tail: T too large");
    MSS_CHECK_CUDA(cudaMemsetAsync(w.zero_base, 0, w.zero_bytes, st));
    if (w.tiles_upper == 0) return MSS_OK;
    // sub-tiles per tile: fat tiles for big inputs (one look-back per 16384 thresholds), >= ~4 tiles per SM for small ones
    const long long T_upper = d_T ? (long long)w.tiles_upper * RC_SUBTILE : (long long)T_host;
    const int sub = (int)std::max<long long>(1, std::min<long long>(RC_SUB_USED, T_upper / ((long long)RC_SUBTILE * 4 * sm_count())));
    const long long tile_thr = (long long)sub * RC_SUBTILE;
    const unsigned tiles = (unsigned)((T_upper + tile_thr - 1) / tile_thr);
    roc_compact_kernel<<<tiles, RC_THREADS + 32, 0, st>>>(tps, fps, d_T, T_host, sub, recall_level, w.status, w.counter,
                                                          w.tile_best, w.tps_k, w.fps_k, w.totals);
    MSS_CHECK_LAUNCH();
    int n_partial = 0;
    if (tiles > 1024) {
        fpr_reduce_kernel<<<256, CT_THREADS, 0, st>>>(w.tile_best, d_T, T_host, tile_thr, w.partial);
        MSS_CHECK_LAUNCH();
        n_partial = 256;
    }
    fpr_final_kernel<<<1, CT_THREADS, 0, st>>>(n_partial ? w.partial : w.tile_best, n_partial, fps, d_T, T_host, tile_thr, w.results);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

// stage 2: T and T_roc are known on the host (they shape numpy's pairwise trees): leaf sums of both integrals in one
// launch, combine below the frontier on the device and above it here.  Synchronises the stream.
static int tail2_run(const int64_t *tps_, const int64_t *fps_, int64_t T, int64_t T_roc, const TailWs &w, cudaStream_t st,
                     double out_host[3]) {
    const long long *tps = (const long long *)tps_, *fps = (const long long *)fps_;
    PwPlan p[2];
    const long long nn[2] = {T, T_roc};
    for (int y = 0; y < 2; y++)
        if (!pw_plan(nn[y], p[y]) || (size_t)p[y].n_leaves > w.max_leaves) {
            set_error("mss_metrics_tail: internal pairwise-plan bound exceeded (n=%lld)", nn[y]);
            return MSS_ERR_WORKSPACE;
        }
    // blob: [ap.size | ap.leaves | roc.size | roc.leaves | ap.front | roc.front], 8-byte fields throughout
    std::vector<char> blob;
    size_t off_size[2], off_leaves[2], off_front[2];
    auto put = [&](const void *src, size_t bytes) { const size_t o = blob.size(); blob.insert(blob.end(), (const char *)src, (const char *)src + bytes); return o; };
    for (int y = 0; y < 2; y++) {
        off_size[y] = put(p[y].size.data(), p[y].size.size() * 8);
        off_leaves[y] = put(p[y].leaves.data(), p[y].leaves.size() * 8);
    }
    for (int y = 0; y < 2; y++) off_front[y] = put(p[y].front.data(), p[y].front.size() * sizeof(PwFrontNode));
    MSS_REQUIRE(blob.size() <= pw_blob_bytes(), "mss_metrics_tail: internal blob bound");
    MSS_CHECK_CUDA(cudaMemcpyAsync(w.blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));

    PwTree2 tt;
    LeafJob lj;
    CombineJob cj;
    double *node_sum[2] = {w.results + 8, w.results + 8 + p[0].front.size()};
    double *leaf_sum[2] = {w.sum_ap, w.sum_roc};
    long long max_leaves = 0;
    int max_front = 0;
    for (int y = 0; y < 2; y++) {
        tt.tree[y] = PwTree{(const long long *)(w.blob + off_size[y]), (const long long *)(w.blob + off_leaves[y]),
                            (int)p[y].size.size(), p[y].n, p[y].n_leaves};
        tt.leaf_start[y] = w.leaf_start[y];
        lj.leaf_start[y] = w.leaf_start[y];
        lj.n_leaves[y] = p[y].n_leaves;
        lj.leaf_sum[y] = leaf_sum[y];
        cj.nodes[y] = (const PwFrontNode *)(w.blob + off_front[y]);
        cj.n_nodes[y] = (int)p[y].front.size();
        cj.leaf_sum[y] = leaf_sum[y];
        cj.node_sum[y] = node_sum[y];
        max_leaves = std::max(max_leaves, p[y].n_leaves);
        max_front = std::max(max_front, (int)p[y].front.size());
    }
    lj.ap = ApTerm{tps, fps, T};
    lj.roc = RocTerm{w.tps_k, w.fps_k, tps + (T - 1), fps + (T - 1)};
    leaf_bounds_kernel<<<dim3((unsigned)((max_leaves + 1 + 255) / 256), 2), 256, 0, st>>>(tt);
    MSS_CHECK_LAUNCH();
    MSS_CHECK_CUDA(cudaFuncSetAttribute(leaf_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LS_SMEM));   // per device: set every time
    leaf_sum_kernel<<<dim3((unsigned)((max_leaves + LS_LEAVES - 1) / LS_LEAVES), 2), LS_THREADS, LS_SMEM, st>>>(lj);
    MSS_CHECK_LAUNCH();
    subtree_combine_kernel<<<dim3((max_front + 127) / 128, 2), 128, 0, st>>>(cj);
    MSS_CHECK_LAUNCH();

    const size_t n_res = 8 + p[0].front.size() + p[1].front.size();
    std::vector<double> h(n_res);
    MSS_CHECK_CUDA(cudaMemcpyAsync(h.data(), w.results, n_res * 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    size_t nx = 0;
    const double ap_sum = pw_combine_top(h.data() + 8, nx, T, p[0].F);
    nx = 0;
    const double auroc = pw_combine_top(h.data() + 8 + p[0].front.size(), nx, T_roc, p[1].F);
    const double ap = -ap_sum;
    out_host[0] = auroc;                  // auc(): direction == 1 because fpr is non-decreasing
    out_host[1] = ap > 0.0 ? ap : 0.0;    // max(0.0, -sum(...))
    out_host[2] = h[0];
    return MSS_OK;
}

}  // namespace mss

using namespace mss;

extern "C" size_t mss_counts_workspace_bytes(int64_t n) { return counts_ws_bytes(n); }

extern "C" int mss_counts_from_sorted(const uint32_t *neg_keys, int64_t n_neg, const uint32_t *pos_keys, int64_t n_pos,
                                      int64_t pos_before, int64_t neg_before, int64_t *tps, int64_t *fps, int64_t *T_host,
                                      void *workspace, size_t workspace_bytes, void *stream) {
    MSS_REQUIRE(n_neg >= 0 && n_pos >= 0 && T_host, "mss_counts_from_sorted: bad arguments");
    *T_host = 0;
    const int64_t n = n_neg + n_pos;
    if (n == 0) return MSS_OK;
    MSS_REQUIRE((n_neg == 0 || neg_keys) && (n_pos == 0 || pos_keys) && tps && fps && workspace, "mss_counts_from_sorted: null pointer");
    CountsWs w;
    if (!carve_counts(workspace, workspace_bytes, n, w)) {
        set_error("mss_counts_from_sorted: workspace too small (%zu < %zu)", workspace_bytes, mss_counts_workspace_bytes(n));
        return MSS_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = plan_enqueue(w.plan, neg_keys, n_neg, pos_keys, n_pos, st);
    if (rc) return rc;
    rc = counts_enqueue(w.plan, n, pos_before, neg_before, tps, fps, w, st);
    if (rc) return rc;
    unsigned long long h = 0;
    MSS_CHECK_CUDA(cudaMemcpyAsync(&h, w.totals, 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    *T_host = (int64_t)h;
    return MSS_OK;
}

extern "C" size_t mss_tail_workspace_bytes(int64_t T) { return tail_ws_bytes(T); }

extern "C" int mss_metrics_tail(const int64_t *tps, const int64_t *fps, int64_t T, double recall_level,
                                void *workspace, size_t workspace_bytes, double out_host[3], int64_t *T_roc_host,
                                void *stream) {
    MSS_REQUIRE(tps && fps && T >= 1 && workspace && out_host, "mss_metrics_tail: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    TailWs w;
    if (!carve_tail(workspace, workspace_bytes, T, w)) {
        set_error("mss_metrics_tail: workspace too small (%zu < %zu)", workspace_bytes, mss_tail_workspace_bytes(T));
        return MSS_ERR_WORKSPACE;
    }
    int rc = tail1_enqueue(tps, fps, nullptr, T, recall_level, w, st);
    if (rc) return rc;
    // T_roc shapes the ROC tree; P, N decide the empty-class outcome
    unsigned long long h_troc = 0;
    long long PN[2];
    MSS_CHECK_CUDA(cudaMemcpyAsync(&h_troc, w.totals, 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(&PN[0], tps + (T - 1), 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(&PN[1], fps + (T - 1), 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    if (T_roc_host) *T_roc_host = (int64_t)h_troc;
    if (PN[0] <= 0 || PN[1] <= 0) {
        set_error("mss_metrics_tail: P=%lld N=%lld (a class is empty)", PN[0], PN[1]);
        return MSS_EMPTY_CLASS;
    }
    return tail2_run(tps, fps, T, (int64_t)h_troc, w, st, out_host);
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
// one-shot layout: [keys n*4][state][sort ws][tps n*8][fps n*8][counts ws][tail ws]
static size_t one_shot_layout(int64_t n, size_t off[7]) {
    size_t o = 0;
    off[0] = o; o += align_up(((size_t)n + 4) * 4, 256);
    off[1] = o; o += 256;
    off[2] = o; o += align_up(sort_ws_bytes(n), 256);
    off[3] = o; o += align_up((size_t)n * 8, 256);
    off[4] = o; o += align_up((size_t)n * 8, 256);
    off[5] = o; o += align_up(counts_ws_bytes(n), 256);
    off[6] = o; o += align_up(tail_ws_bytes(n), 256);
    return o;
}

extern "C" size_t mss_ood_metrics_workspace_bytes(int64_t n) {
    size_t off[7];
    return one_shot_layout(n < 0 ? 0 : n, off) + 256;
}

static int check_state(const EvalState &h) {
    // the reference checks emptiness first (metric.py:176), sklearn validates finiteness before anything else it does
    if (h.n_pos == 0 || h.n_neg == 0) return MSS_EMPTY_CLASS;
    if (h.nan_flag) { set_error("Input contains NaN."); return MSS_ERR_NAN; }
    if (h.inf_flag) { set_error("Input contains infinity or a value too large for dtype('float32')."); return MSS_ERR_INF; }
    return MSS_OK;
}

// Everything up to the point where the host must know T / T_roc is enqueued without a round trip: the stream sizes
// are read by the kernels from the evaluator state, grids are sized for n_upper.  ONE synchronisation fetches the
// state and the two totals, a second one the leaf sums.
static int metrics_speculative(const mss_eval_buffers *ev, int64_t n_upper, char *ws_sort, size_t sort_b, int64_t *tps,
                               int64_t *fps, char *ws_counts, size_t counts_b, char *ws_tail, size_t tail_b,
                               double out_host[3], int64_t counts_host[4], cudaStream_t st) {
    const SortPlan *plan = nullptr;
    int rc = sort_enqueue(ev, nullptr, 0, nullptr, 0, n_upper, ws_sort, sort_b, st, &plan);
    if (rc) return rc;
    CountsWs cw;
    TailWs tw;
    if (!carve_counts(ws_counts, counts_b, n_upper, cw) || !carve_tail(ws_tail, tail_b, n_upper, tw)) {
        set_error("mss_ood_metrics: workspace too small");
        return MSS_ERR_WORKSPACE;
    }
    rc = counts_enqueue(plan, n_upper, 0, 0, tps, fps, cw, st);
    if (rc) return rc;
    rc = tail1_enqueue(tps, fps, cw.totals, 0, 0.95, tw, st);
    if (rc) return rc;
    EvalState h;
    unsigned long long h_T = 0, h_Troc = 0;
    MSS_CHECK_CUDA(cudaMemcpyAsync(&h, ev->state, sizeof(h), cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(&h_T, cw.totals, 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(&h_Troc, tw.totals, 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    if (h.overflow || (int64_t)(h.n_neg + h.n_pos) > ev->capacity) {
        set_error("evaluator capacity %lld exceeded (%llu valid pixels appended)", (long long)ev->capacity,
                  (unsigned long long)(h.n_neg + h.n_pos + h.overflow));
        return MSS_ERR_WORKSPACE;
    }
    rc = check_state(h);
    if (rc) return rc;
    if (counts_host) { counts_host[0] = (int64_t)h.n_pos; counts_host[1] = (int64_t)h.n_neg; counts_host[2] = (int64_t)h_T; counts_host[3] = (int64_t)h_Troc; }
    return tail2_run(tps, fps, (int64_t)h_T, (int64_t)h_Troc, tw, st, out_host);
}

extern "C" int mss_ood_metrics(const float *scores, const void *labels, int label_dtype, int64_t n, int64_t id_in,
                               int64_t id_out, void *workspace, size_t workspace_bytes, double out_host[3],
                               int64_t counts_host[4], void *stream) {
    MSS_REQUIRE(n >= 0 && out_host, "mss_ood_metrics: bad arguments");
    if (n == 0) return MSS_EMPTY_CLASS;
    MSS_REQUIRE(scores && labels && workspace, "mss_ood_metrics: null pointer");
    size_t off[7];
    const size_t need = one_shot_layout(n, off);
    char *ws = (char *)align_up((size_t)(uintptr_t)workspace, 256);
    if ((size_t)(ws - (char *)workspace) + need > workspace_bytes) {
        set_error("mss_ood_metrics: workspace too small (%zu < %zu)", workspace_bytes, mss_ood_metrics_workspace_bytes(n));
        return MSS_ERR_WORKSPACE;
    }
    mss_eval_buffers ev{(uint32_t *)(ws + off[0]), ws + off[1], n};
    int rc = mss_eval_reset(&ev, stream);
    if (rc) return rc;
    rc = mss_eval_append(scores, labels, label_dtype, n, id_in, id_out, &ev, stream);
    if (rc) return rc;
    return metrics_speculative(&ev, n, ws + off[2], off[3] - off[2], (int64_t *)(ws + off[3]), (int64_t *)(ws + off[4]),
                               ws + off[5], off[6] - off[5], ws + off[6], need - off[6], out_host, counts_host,
                               (cudaStream_t)stream);
}

// From an evaluator: the host reads the state first (it also sizes the workspace), so every stage gets exact sizes.
//   workspace >= mss_ood_metrics_from_eval_workspace_bytes(count)   with count = mss_eval_state_host()[0]
extern "C" size_t mss_ood_metrics_from_eval_workspace_bytes(int64_t count) {
    if (count < 0) count = 0;
    return 256 + align_up(sort_ws_bytes(count), 256) + 2 * align_up((size_t)count * 8, 256) + align_up(counts_ws_bytes(count), 256) +
           align_up(tail_ws_bytes(count), 256);
}

extern "C" int mss_ood_metrics_from_eval(const mss_eval_buffers *ev, void *workspace, size_t workspace_bytes,
                                         double out_host[3], int64_t counts_host[4], void *stream) {
    MSS_REQUIRE(ev && ev->keys && ev->state && workspace && out_host, "mss_ood_metrics_from_eval: null pointer");
    int64_t s4[4];
    int rc = mss_eval_state_host(ev, s4, stream);
    if (rc) return rc;
    const int64_t m = s4[0];
    if (s4[1] == 0 || s4[1] == m) return MSS_EMPTY_CLASS;
    if (s4[2]) { set_error("Input contains NaN."); return MSS_ERR_NAN; }
    if (s4[3]) { set_error("Input contains infinity or a value too large for dtype('float32')."); return MSS_ERR_INF; }
    char *ws = (char *)align_up((size_t)(uintptr_t)workspace, 256);
    const size_t lead = (size_t)(ws - (char *)workspace);
    const size_t sort_b = align_up(sort_ws_bytes(m), 256), cnt_b = align_up((size_t)m * 8, 256);
    const size_t cw_b = align_up(counts_ws_bytes(m), 256), tw_b = align_up(tail_ws_bytes(m), 256);
    if (lead + sort_b + 2 * cnt_b + cw_b + tw_b > workspace_bytes) {
        set_error("mss_ood_metrics_from_eval: workspace too small (%zu < %zu)", workspace_bytes, lead + sort_b + 2 * cnt_b + cw_b + tw_b);
        return MSS_ERR_WORKSPACE;
    }
    return metrics_speculative(ev, m, ws, sort_b, (int64_t *)(ws + sort_b), (int64_t *)(ws + sort_b + cnt_b), ws + sort_b + 2 * cnt_b,
                               cw_b, ws + sort_b + 2 * cnt_b + cw_b, tw_b, out_host, counts_host, (cudaStream_t)stream);
}

/* stage-level: sort both streams of an evaluator in place (the multi-GPU evaluator sorts, counts and runs the tail as
 * separate steps around its collectives).  n_upper >= the number of stored keys. */
extern "C" int mss_eval_sort(const mss_eval_buffers *ev, int64_t n_upper, void *workspace, size_t workspace_bytes, void *stream) {
    MSS_REQUIRE(ev && ev->keys && ev->state && workspace && n_upper >= 0, "mss_eval_sort: bad arguments");
    return sort_enqueue(ev, nullptr, 0, nullptr, 0, n_upper, workspace, workspace_bytes, (cudaStream_t)stream, nullptr);
}
