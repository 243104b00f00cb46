// Stable LSD radix sort of uint32 keys -- the GPU replacement for the three full argsorts the reference
// performs per evaluation (sklearn _ranking.py:909 inside roc_auc_score and again inside
// average_precision_score, plus np.argsort(kind="mergesort") at lib/utils/metric.py:103).
//
// Round 2: KEY-ONLY and SEGMENTED.  The evaluator keeps in-distribution and OOD keys in two streams (eval_append.cuh),
// so the label is never carried through the sort: 4 + 4 x (4 + 4) = 36 B of HBM traffic per key instead of the 44 B per
// (key, u8 label) pair of round 1, no 1-byte scatter (ncu, round 1: 2.2x store-sector inflation), and no label staging
// in shared memory.  Both streams are sorted by ONE sequence of launches: tiles are numbered over both segments and a
// tile's look-back stops at the first tile of its own segment.
//
// "Onesweep": one upfront histogram of all four 8-bit digits, then four scatter passes, each reading and writing every
// key exactly once; the cross-tile digit offsets come from a decoupled look-back over per-tile status words instead of
// a separate scan pass.  A pass whose digit is the same for every key of a segment (fp16-born or quantised scores
// leave the low byte constant) is skipped for that segment; the decision is taken on the device from the histogram.
// Tile = 256 threads x 16 keys.  Ranking inside a tile is stable and atomics-free: keys are warp-striped, each warp
// ranks its 512 keys item by item with ballots, warps are combined by a per-digit prefix over the 8 warps.
//
// The same tile routine, with the digit replaced by "which key range does this key fall in" (SplitterDigit), is the
// local half of the multi-GPU key-range exchange: bucket d is stored straight into rank d's receive buffer
// (mss_partition_scatter_keys).
#include <stdlib.h>

#include "sort_plan.cuh"

namespace mss {

constexpr int RADIX = 256;
constexpr int SORT_THREADS = 256;
constexpr int SORT_IPT = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;  // 4096
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int LB_WINDOW = 8;                        // look-back statuses fetched per round trip (16: 39.7 vs 40.8 Gpairs/s)

struct ShiftDigit {
    static constexpr bool kCheap = true;
    int shift;
    __device__ __forceinline__ unsigned operator()(uint32_t k) const { return (k >> shift) & 255u; }
};

// dest(key) = #{ j : key >= splitter[j] }  (splitters ascending, nspl = parts - 1 <= 255 of them).
// `spl` points at a SHARED-memory copy inside the kernels; branchless lower bound over 2^steps - 1 slots, the slots past
// nspl acting as +infinity.
struct SplitterDigit {
    static constexpr bool kCheap = false;
    const uint32_t *spl;
    int nspl, steps;
    __device__ __forceinline__ unsigned operator()(uint32_t k) const {
        unsigned lo = 0;
        for (int s = 1 << (steps - 1); s > 0; s >>= 1) {
            const unsigned idx = lo + s - 1;
            lo += ((int)idx < nspl && k >= spl[idx]) ? s : 0;
        }
        return lo;
    }
};
static int splitter_steps(int parts) {
    int steps = 1;
    while ((1 << steps) - 1 < parts - 1) steps++;
    return steps;
}

// Lanes holding the same BITS-bit value as this lane, from BITS ballots.  (match.any does the same in one
// instruction but costs ~200 cycles per warp on sm_100 when the 32 values are distinct -- measured with
// ncu in round 1: it made the histogram 9x and the scatter pass 2x slower than this form.)
template <int BITS>
__device__ __forceinline__ unsigned match_bits(unsigned d, bool valid) {
    unsigned peers = __ballot_sync(0xffffffffu, valid);
    if (!valid) peers = ~peers;                  // lanes past the end group among themselves
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        const bool bit = (d >> b) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
}

// ---- the plan -------------------------------------------------------------------------------------------
__global__ void sort_plan_kernel(SortPlan *plan, uint32_t *xa, uint32_t *ya, long long na, uint32_t *xb, uint32_t *yb,
                                 long long nb, const EvalState *state, uint32_t *keys, uint32_t *alt, long long capacity) {
    if (state) {
        // evaluator streams: negatives at keys[0, n_neg), positives at keys[capacity - n_pos, capacity)
        na = (long long)min(state->n_neg, (unsigned long long)capacity);
        nb = (long long)min(state->n_pos, (unsigned long long)(capacity - na));
        xa = keys;
        xb = keys + (capacity - nb);
        ya = alt;
        yb = alt + ((na + 3) & ~3ll);
    }
    SortPlan p;
    p.seg[0] = SortSeg{xa, ya, na, 0u, (unsigned)((na + SORT_TILE - 1) / SORT_TILE)};
    p.seg[1] = SortSeg{xb, yb, nb, p.seg[0].tiles, (unsigned)((nb + SORT_TILE - 1) / SORT_TILE)};
    p.total_tiles = p.seg[0].tiles + p.seg[1].tiles;
    for (int s = 0; s < 2; s++) {
        for (int q = 0; q < 4; q++) p.sel[s][q] = (unsigned char)(q & 1);
        p.copy_back[s] = 0;
    }
    p.pad = 0;
    *plan = p;
}

// ---- upfront histogram of the four digits ------------------------------------------------------------
// Shared-memory atomics cost ~1-2 cycles per LANE on sm_100 and the digits of real score distributions are
// extremely skewed (sign/exponent bits; zeroed mantissa bits of fp16-born scores), so contention-sensitive
// schemes are out.  Every LANE owns private counters instead -- plain load / add / store, no atomics, no
// dependence on the key distribution.  History (ncu, B200):
//   round 1, first form  4 warps per SM, 32-bit counters, register-prefetched global loads: 814 us per 32 M keys, issue
//                        6.5 %, 4 of 64 warp slots -- pure latency;
//   round 1, second form 16-bit counters (16 KB per warp) -> 12 warps per SM, keys through a 4-stage ring of bulk copies:
//                        394 us per 134 M keys; 40 % of the stall samples on the shared-memory round trip of the
//                        load-add-store chain (short scoreboard), issue 47 %;
//   round 2, this form   8-BIT counters (8 KB per warp) -> 20 warps per SM to hide that round trip: warp = (group g of 5,
//                        digit d of 4); the four digit-warps of a group count the same keys, the five groups split every
//                        chunk; counter (bin b, lane l) of a warp is the byte at b*32 + l.  Measured: 448 us per 134 M
//                        keys INCLUDING the zeroing of the passes' 268 MB of status words (side job below) -- the same as
//                        the second form plus its separate memset; the kernel is bound by its ~52 instructions per key
//                        (IMAD / shift address arithmetic of four digits), not by the round trip any more.
//   * keys arrive through a 4-stage ring of 10 KB bulk copies (cp.async.bulk + mbarrier, one elected lane);
//   * a lane adds at most 16 to one counter per chunk, so the bytes are folded into 32-bit per-CTA totals every
//     HIST_EPOCH = 15 chunks (<= 240 per byte) -- warp-local, no CTA barrier; the fold sums a bin's 32 bytes with 8 dp4a.
// Two keys are in flight per lane; when they hit the same counter both store the merged total.
// blockIdx.y = segment of the plan; a segment uses as many of the gridDim.x CTAs as it has work for.
constexpr int HIST_GROUPS = 5;
constexpr int HIST_WARPS = HIST_GROUPS * 4;                    // 20
constexpr int HIST_THREADS = HIST_WARPS * 32;                  // 640
constexpr int HIST_CHUNK_KEYS = HIST_GROUPS * 32 * 4 * 4;      // 2560 keys = 10 KB: 4 uint4 per lane per group
constexpr int HIST_CHUNK_BYTES = HIST_CHUNK_KEYS * 4;
constexpr int HIST_STAGES = 4;
constexpr int HIST_WARP_BYTES = RADIX * 32;                    // 8 KB of u8 counters
constexpr int HIST_EPOCH = 15;
constexpr int HIST_SMEM = HIST_WARPS * HIST_WARP_BYTES + HIST_STAGES * HIST_CHUNK_BYTES + 4 * RADIX * 4 +
                          2 * HIST_STAGES * 8 + 128;

__device__ __forceinline__ unsigned lds_u8(uint32_t a) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, unsigned v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

__global__ void __launch_bounds__(HIST_THREADS, 1)
radix_histogram_kernel(const SortPlan *__restrict__ plan, unsigned long long *__restrict__ hist_all, int epoch_chunks,
                       uint4 *__restrict__ zero_base, size_t zero_chunks) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    // Side job: zero the look-back status words of the four passes (2 KB per tile and pass: 268 MB for 134 M keys).
    // This kernel is bound by its shared-memory counter updates, its store path is idle.
    {
        const size_t nthr = (size_t)gridDim.x * gridDim.y * HIST_THREADS;
        for (size_t i = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * HIST_THREADS + threadIdx.x; i < zero_chunks; i += nthr)
            zero_base[i] = make_uint4(0, 0, 0, 0);
    }
    const SortSeg seg = plan->seg[blockIdx.y];
    const uint32_t *__restrict__ keys = seg.x;
    const long long n = seg.n;
    unsigned long long *__restrict__ hist = hist_all + blockIdx.y * 4 * RADIX;
    // CTAs this segment employs: >= 4 chunks per CTA
    const unsigned grid = (unsigned)max(1ll, min((long long)gridDim.x, (n + 4 * HIST_CHUNK_KEYS - 1) / (4 * HIST_CHUNK_KEYS)));
    if (n == 0 || blockIdx.x >= grid) return;

    unsigned char *s_cnt = s_raw;                                                  // [20][256][32] u8
    uint4 *s_ring = reinterpret_cast<uint4 *>(s_raw + HIST_WARPS * HIST_WARP_BYTES);   // [4][640] uint4
    unsigned *s_tot = reinterpret_cast<unsigned *>(s_ring + HIST_STAGES * (HIST_CHUNK_KEYS / 4));   // [4][256]
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_tot + 4 * RADIX), *s_empty = s_full + HIST_STAGES;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned digit = warp & 3, group = warp >> 2;

    for (int i = tid; i < HIST_WARPS * HIST_WARP_BYTES / 16; i += HIST_THREADS)
        reinterpret_cast<uint4 *>(s_cnt)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 4 * RADIX; i += HIST_THREADS) s_tot[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < HIST_STAGES; s++) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], HIST_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // keys before the first 16-byte boundary (head) and after the last whole uint4 (tail): CTA 0, below.
    const long long head = min(n, (long long)(((16 - ((uintptr_t)keys & 15)) & 15) >> 2));
    const long long groups4 = (n - head) >> 2;                                     // whole uint4 groups
    const uint4 *k4 = reinterpret_cast<const uint4 *>(keys + head);
    const long long chunks = (groups4 + HIST_CHUNK_KEYS / 4 - 1) / (HIST_CHUNK_KEYS / 4);
    // CTA c owns chunks c, c + grid, ...
    const long long my_chunks = (chunks > (long long)blockIdx.x) ? (chunks - blockIdx.x + grid - 1) / grid : 0;
    auto issue = [&](long long j) {                                                // elected lane only
        const long long c = (long long)blockIdx.x + j * grid;
        const long long g0 = c * (HIST_CHUNK_KEYS / 4);
        const unsigned bytes = (unsigned)(min((long long)(HIST_CHUNK_KEYS / 4), groups4 - g0) * 16);
        const int s = (int)(j % HIST_STAGES);
        mbar_expect_tx(&s_full[s], bytes);
        bulk_load_1d(s_ring + s * (HIST_CHUNK_KEYS / 4), k4 + g0, bytes, &s_full[s]);
    };
    if (tid == 0)
        for (long long j = 0; j < HIST_STAGES && j < my_chunks; j++) issue(j);

    const uint32_t mine = smem_u32(s_cnt) + warp * HIST_WARP_BYTES + lane;
    const int shift = 8 * (int)digit;
    // fold this warp's 8-bit counters into the CTA totals of its digit and zero them (warp-local)
    auto fold = [&]() {
        __syncwarp();
        uint32_t *w32 = reinterpret_cast<uint32_t *>(s_cnt + warp * HIST_WARP_BYTES);
#pragma unroll 1
        for (int r = 0; r < RADIX / 32; r++) {
            const int bin = r * 32 + lane;
            unsigned acc = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {                                          // rotated: 32 lanes, 32 banks
                const int wi = bin * 8 + ((j + (lane >> 2)) & 7);
                acc = __dp4a(w32[wi], 0x01010101u, acc);
                w32[wi] = 0;
            }
            if (acc) atomicAdd(&s_tot[digit * RADIX + bin], acc);
        }
        __syncwarp();
    };

    int in_epoch = 0;
    for (long long j = 0; j < my_chunks; j++) {
        const int s = (int)(j % HIST_STAGES);
        const unsigned parity = (unsigned)((j / HIST_STAGES) & 1);
        mbar_wait(&s_full[s], parity);
        const long long c = (long long)blockIdx.x + j * grid;
        const int valid4 = (int)min((long long)(HIST_CHUNK_KEYS / 4), groups4 - c * (HIST_CHUNK_KEYS / 4));
        const uint4 *src = s_ring + s * (HIST_CHUNK_KEYS / 4) + group * 128 + lane;
#pragma unroll
        for (int it = 0; it < 4; it++) {
            if ((int)(group * 128 + it * 32 + lane) < valid4) {
                const uint4 v = src[it * 32];
                const uint32_t k[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int h = 0; h < 4; h += 2) {
                    const uint32_t a0 = mine + (((k[h] >> shift) & 255u) << 5);
                    const uint32_t a1 = mine + (((k[h + 1] >> shift) & 255u) << 5);
                    const unsigned add = 1u + (a0 == a1);
                    const unsigned c0 = lds_u8(a0), c1 = lds_u8(a1);
                    sts_u8(a0, c0 + add);
                    sts_u8(a1, c1 + add);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_one(&s_empty[s]);
        if (tid == 0 && j + HIST_STAGES < my_chunks) {                             // refill the stage just drained
            mbar_wait(&s_empty[s], parity);
            issue(j + HIST_STAGES);
        }
        if (++in_epoch == epoch_chunks) { fold(); in_epoch = 0; }
    }
    fold();
    __syncthreads();
    for (int i = tid; i < 4 * RADIX; i += HIST_THREADS)
        if (s_tot[i]) atomicAdd(hist + i, (unsigned long long)s_tot[i]);
    if (blockIdx.x == 0 && tid == 0)                             // <= 3 head + <= 3 tail keys
        for (long long i = 0; i < n; i++) {
            if (i == head) i += groups4 << 2;
            if (i >= n) break;
            const uint32_t kk = __ldg(keys + i);
#pragma unroll
            for (int d = 0; d < 4; d++) atomicAdd(hist + d * RADIX + ((kk >> (8 * d)) & 255u), 1ull);
        }
}

// hist[seg][pass][256] -> exclusive prefix (in place): warp w scans the histogram of (segment w / 4, pass w % 4), 8 bins
// per lane.  A digit whose histogram has a single non-empty bin leaves every key where it is: the pass is skipped for
// that segment (allow_skip).
__global__ void __launch_bounds__(RADIX)
radix_scan_bins_kernel(unsigned long long *__restrict__ hist, SortPlan *plan, int allow_skip) {
    __shared__ unsigned char s_skip[8];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long n = (unsigned long long)plan->seg[warp >> 2].n;
    unsigned long long *h = hist + warp * RADIX + lane * 8;
    unsigned long long v[8], sum = 0;
    bool single = false;
#pragma unroll
    for (int j = 0; j < 8; j++) { v[j] = h[j]; sum += v[j]; single |= (v[j] == n); }
    unsigned long long inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    unsigned long long run = inc - sum;
#pragma unroll
    for (int j = 0; j < 8; j++) { h[j] = run; run += v[j]; }
    const bool any_single = __any_sync(0xffffffffu, single);
    if (lane == 0) s_skip[warp] = (unsigned char)((allow_skip && n > 0 && any_single) || n == 0);
    __syncthreads();
    if (threadIdx.x < 2) {
        const int s = threadIdx.x;
        int cur = 0;                                         // 0: data in x, 1: data in y
        for (int p = 0; p < 4; p++) {
            if (s_skip[s * 4 + p]) plan->sel[s][p] = 2;
            else { plan->sel[s][p] = (unsigned char)cur; cur ^= 1; }
        }
        plan->copy_back[s] = (unsigned char)cur;
    }
}

// ---- one scatter pass -----------------------------------------------------------------------------
// tile_status[tile][digit]: look-back status words (common.cuh)

// 8-bit match for a tile without padding lanes: 4 instructions per bit (predicate, ballot, select, and)
// Written in PTX so that one bit costs exactly predicate (and + setp -> one LOP3 with a predicate result), VOTE, SEL,
// LOP3; the C form `bit = (d >> b) & 1; peers &= bit ? m : ~m` compiled to 6 instructions per bit (shift, and,
// compare, select, vote, lop3) -- 48 of the pass's ~120 instructions per key (SASS, round 1).
template <int B>
__device__ __forceinline__ void match_bit(unsigned &peers, unsigned d) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 t, m, x;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
        "selp.b32 x, 0, 0xffffffff, p;\n\t"
        "lop3.b32 %0, %0, m, x, 0x60;\n\t}"                // peers & (m ^ x)
        : "+r"(peers)
        : "r"(d), "n"(1u << B));
}
__device__ __forceinline__ unsigned match8_full(unsigned d) {
    unsigned peers = 0xffffffffu;
    match_bit<0>(peers, d); match_bit<1>(peers, d); match_bit<2>(peers, d); match_bit<3>(peers, d);
    match_bit<4>(peers, d); match_bit<5>(peers, d); match_bit<6>(peers, d); match_bit<7>(peers, d);
    return peers;
}

// DIG: the digit function is expensive (SplitterDigit): digits are computed once, kept packed in registers and
// staged next to the keys for the write-out; the splitters sit in shared memory.
template <bool DIG>
struct SweepSmem {
    __align__(16) uint32_t keys[SORT_TILE];
    unsigned warp_hist[SORT_WARPS][RADIX];   // per-warp digit counts -> exclusive scatter bases
    unsigned long long global[RADIX];        // byte address of tile-sorted position 0 of each digit in the output
    unsigned long long dst[RADIX];           // byte address where the first key with digit d of the whole segment goes
    unsigned long long *status;              // status words of this segment's tile 0
    unsigned scan[SORT_WARPS];
    // tile context (thread 0 -> everyone)
    const uint32_t *in;
    unsigned long long out;                  // byte address of the output array (sort passes)
    long long n;
    unsigned tile, seg;
    int skip;
    uint8_t dig[DIG ? SORT_TILE : 16];
    uint32_t spl[DIG ? RADIX : 4];
};

// One tile: rank, scatter into tile-sorted order in shared memory, look back, write out.
//   sm.dst[d]  byte address where the FIRST key with digit d of the whole segment goes (written by thread d before)
//   sm.status  status words of this segment's tile 0, `sstride` words per tile; nd live digits (threads d >= nd idle)
// (both parked in shared memory: as arguments they stayed live in registers across the whole tile and the 64-register
// kernel spilled)
// FULL: the tile holds exactly SORT_TILE keys (every tile but possibly the last): no bounds predicates.
template <typename DigitFn, bool FULL, bool DIG>
__device__ __forceinline__ void sweep_tile(SweepSmem<DIG> &sm, int sstride, int nd, DigitFn digit_of, int match_nbits = 8) {
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // (tile context is re-read from shared memory where it is needed instead of being carried in registers)
    const int tile_n = FULL ? SORT_TILE : (int)(sm.n - (long long)sm.tile * SORT_TILE);

    // keys, warp-striped: item i of lane l in warp w sits at w*512 + i*32 + l
    uint32_t key[SORT_IPT];
    const int wbase = warp * (32 * SORT_IPT) + lane;
    const uint32_t *kp = sm.in + (long long)sm.tile * SORT_TILE + wbase;
#pragma unroll
    for (int i = 0; i < SORT_IPT; i++) key[i] = (FULL || wbase + i * 32 < tile_n) ? __ldg(kp + i * 32) : 0u;

    unsigned dpk[DIG ? SORT_IPT / 4 : 1];
    if (DIG) {
#pragma unroll
        for (int i = 0; i < SORT_IPT; i++) {
            const unsigned d = digit_of(key[i]);
            dpk[i >> 2] = (i & 3) ? (dpk[i >> 2] | (d << (8 * (i & 3)))) : d;
        }
    }
    auto dig = [&](int i) -> unsigned {
        if (DIG) return (dpk[i >> 2] >> (8 * (i & 3))) & 255u;
        return digit_of(key[i]);
    };

    // phase 1 (independent, pipelined): lanes of my warp holding the same digit as my item i
    unsigned peers[SORT_IPT];
#pragma unroll
    for (int i = 0; i < SORT_IPT; i++) {
        if (DIG) {
            // destinations fit in digit_of.steps bits (8 GPUs: 3 ballots per key instead of 8)
            const bool valid = FULL || wbase + i * 32 < tile_n;
            const unsigned d = valid ? dig(i) : 0u;
            unsigned pm = __ballot_sync(0xffffffffu, valid);
            if (!valid) pm = ~pm;
            for (int b = 0; b < match_nbits; b++) {
                const bool bit = (d >> b) & 1u;
                const unsigned mm = __ballot_sync(0xffffffffu, bit);
                pm &= bit ? mm : ~mm;
            }
            peers[i] = pm;
        } else if (FULL) peers[i] = match8_full(dig(i));
        else {
            const bool valid = wbase + i * 32 < tile_n;
            peers[i] = match_bits<8>(valid ? dig(i) : 0u, valid);
        }
    }
    // phase 2 (serial per warp; item order == memory order, hence stable): running per-digit counters.
    // two 16-bit ranks per register (8 registers, not 16: the kernel has to fit 64 registers for 4 CTAs per SM)
    unsigned rank2[SORT_IPT / 2];
    const unsigned lt = lanemask_lt();
    unsigned *wh = sm.warp_hist[warp];
#pragma unroll
    for (int i = 0; i < SORT_IPT; i++) {
        const bool valid = FULL || (wbase + i * 32 < tile_n);
        const unsigned below = __popc(peers[i] & lt);
        unsigned old = 0;
        if (below == 0 && valid) {
            const unsigned d = dig(i);
            old = wh[d];
            wh[d] = old + __popc(peers[i]);
        }
        old = __shfl_sync(0xffffffffu, old, __ffs(peers[i]) - 1);
        const unsigned r16 = old + below;
        rank2[i >> 1] = (i & 1) ? (rank2[i >> 1] | (r16 << 16)) : r16;
        __syncwarp();
    }
    __syncthreads();

    // thread d: prefix over the 8 warps for digit d, tile count, publish
    {
        const unsigned d = tid, tile = sm.tile;
        unsigned cnt[SORT_WARPS], tot = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) { cnt[w] = sm.warp_hist[w][d]; tot += cnt[w]; }
        if ((int)d < nd) st_status(sm.status + (size_t)tile * sstride + d, (tile == 0 ? FLAG_INC : FLAG_AGG) | tot);

        // exclusive scan of tot over the 256 digits -> first tile-sorted position of digit d
        unsigned inc = tot;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, inc, s);
            if (lane >= s) inc += t;
        }
        if (lane == 31) sm.scan[warp] = inc;
        __syncthreads();
        unsigned start = inc - tot;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) start += (w < (int)warp) ? sm.scan[w] : 0u;
        // scatter base of (warp w, digit d) = start + counts of the warps before w
        unsigned run = start;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) { sm.warp_hist[w][d] = run; run += cnt[w]; }
        sm.global[d] = (unsigned long long)tot | ((unsigned long long)start << 32);   // parked until the look-back below
    }
    __syncthreads();

    // scatter into tile-sorted order in smem.  This needs only tile-local offsets, so it runs BEFORE the look-back: the
    // predecessors get this much more time to publish their inclusive prefixes and the look-back below finds one after a
    // window or two (ncu, round 1: look-back straight after the count walked ~26 tiles back and was 25 % of all
    // instructions).
#pragma unroll
    for (int i = 0; i < SORT_IPT; i++) {
        if (FULL || wbase + i * 32 < tile_n) {
            const unsigned r16 = (i & 1) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xffffu);
            const unsigned d = dig(i);
            const unsigned pos = wh[d] + r16;
            sm.keys[pos] = key[i];
            if (DIG) sm.dig[pos] = (uint8_t)d;
        }
    }

    // Decoupled look-back, thread d for digit d, LB_WINDOW predecessors per round trip (the loads of one window
    // are independent, so they overlap); branch-light: one "all published?" test per window, predicated adds.
    if ((int)tid < nd) {
        const unsigned d = tid, tile = sm.tile;
        const unsigned tot = (unsigned)sm.global[d], start = (unsigned)(sm.global[d] >> 32);
        unsigned long long *my = sm.status + (size_t)tile * sstride + d;
        unsigned long long prefix = 0;
        if (tile > 0) {
            int t = (int)tile - 1;                                  // tiles < 2^31 (host-checked)
            const unsigned long long *p = my - sstride;
            for (;;) {
                unsigned long long st[LB_WINDOW];
                if (t >= LB_WINDOW - 1) {                           // CTA-uniform: all predecessors of the window exist
#pragma unroll
                    for (int j = 0; j < LB_WINDOW; j++) st[j] = ld_status(p - (size_t)j * sstride);
                } else {
#pragma unroll
                    for (int j = 0; j < LB_WINDOW; j++) st[j] = (j <= t) ? ld_status(p - (size_t)j * sstride) : FLAG_INC;   // before tile 0: prefix 0
                }
                // flags live in the top two bits: 32-bit tests on the high words
                unsigned lowest = 0xffffffffu;
#pragma unroll
                for (int j = 0; j < LB_WINDOW; j++) lowest = min(lowest, (unsigned)(st[j] >> 32));
                if (lowest < 0x40000000u) continue;                 // some predecessor has not even counted yet: read again
                bool done = false;                                  // an inclusive prefix has been met
#pragma unroll
                for (int j = 0; j < LB_WINDOW; j++) {
                    if (!done) prefix += st[j] & VAL_MASK;
                    done = done || (unsigned)(st[j] >> 32) >= 0x80000000u;
                }
                if (done) break;
                t -= LB_WINDOW;
                p -= (size_t)LB_WINDOW * sstride;
            }
            st_status(my, FLAG_INC | (prefix + tot));
        }
        // byte address of tile-sorted position 0 "as if" it belonged to digit d's run
        sm.global[d] = sm.dst[d] + 4ull * prefix - 4ull * start;  // (only thread d ever touches sm.global[d] / sm.dst[d] up to here)
    }
    __syncthreads();

    // coalesced write-out: equal digits are contiguous in smem and in the output
#pragma unroll
    for (int i = 0; i < SORT_IPT; i++) {
        const int p = i * SORT_THREADS + tid;
        if (FULL || p < tile_n) {
            const uint32_t k = sm.keys[p];
            const unsigned d = DIG ? (unsigned)sm.dig[p] : digit_of(k);
            *reinterpret_cast<uint32_t *>(sm.global[d] + 4ull * (unsigned)p) = k;
        }
    }
}

#ifndef SORT_CTAS_PER_SM
#define SORT_CTAS_PER_SM 4
#endif
// One pass over both segments of the plan.  Tiles are numbered in START order (atomic ticket), so the look-back only
// ever waits on tiles that are already running.
__global__ void __launch_bounds__(SORT_THREADS, SORT_CTAS_PER_SM)
onesweep_pass_kernel(const SortPlan *__restrict__ plan, int pass, const unsigned long long *__restrict__ hist_all,
                     unsigned long long *tile_status, unsigned *tile_counter) {
    __shared__ SweepSmem<false> sm;
    const unsigned tid = threadIdx.x;
    if (tid == 0) {
        const unsigned t = atomicAdd(tile_counter, 1u);
        const unsigned s = t >= plan->seg[1].tile0 ? 1u : 0u;
        const SortSeg sg = plan->seg[s];
        const unsigned sel = plan->sel[s][pass];
        sm.tile = t - sg.tile0;
        sm.seg = s;
        sm.skip = (t >= plan->total_tiles) || sel == 2;
        sm.in = sel == 0 ? sg.x : sg.y;
        sm.out = (unsigned long long)(uintptr_t)(sel == 0 ? sg.y : sg.x);
        sm.n = sg.n;
        sm.status = tile_status + (size_t)sg.tile0 * RADIX;
    }
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) sm.warp_hist[w][tid] = 0;
    __syncthreads();
    if (sm.skip) return;
    sm.dst[tid] = sm.out + 4ull * __ldg(hist_all + (sm.seg * 4 + pass) * RADIX + tid);
    const ShiftDigit dg{8 * pass};
    if ((long long)(sm.tile + 1) * SORT_TILE <= sm.n)
        sweep_tile<ShiftDigit, true, false>(sm, RADIX, RADIX, dg);
    else
        sweep_tile<ShiftDigit, false, false>(sm, RADIX, RADIX, dg);
}

// (Round-2 experiment, not kept: a PERSISTENT form of this kernel -- 4 resident CTAs per SM walking tickets, the next
// tile's keys loaded into the key registers right after the shared-memory scatter so that they arrive during the
// look-back and the write-out -- was no faster: 44.2 vs 44.7 Gkeys/s at 134 M keys, 32.2 vs 35.0 at 8 M.  With four
// independent CTAs per SM the hardware already overlaps one tile's load latency with the others' arithmetic; the pass is
// bound by its ~84 instructions per key (ALU pipe 58 %, issue 53 %), not by exposed memory latency.)

// a segment whose last executed pass wrote into the alternate buffer is copied back (odd number of live passes)
__global__ void __launch_bounds__(256)
sort_copy_back_kernel(const SortPlan *__restrict__ plan) {
    for (int s = 0; s < 2; s++) {
        if (!plan->copy_back[s]) continue;
        const SortSeg sg = plan->seg[s];
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < sg.n; i += (long long)gridDim.x * blockDim.x)
            sg.x[i] = sg.y[i];
    }
}

// Multi-GPU exchange, local half: stable partition by key range where bucket d is stored at byte address table[d]
// (+ this rank's running offset inside the bucket) -- typically inside rank d's peer-mapped receive buffer, so the
// stores travel over NVLink and neither a staging copy nor an all-to-all is needed.
__global__ void __launch_bounds__(SORT_THREADS, 3)
partition_scatter_kernel(const uint32_t *__restrict__ keys, long long n, const uint32_t *__restrict__ splitters, int nspl,
                         int steps, const unsigned long long *__restrict__ table, unsigned long long *tile_status,
                         int sstride, unsigned *tile_counter) {
    __shared__ SweepSmem<true> sm;
    const unsigned tid = threadIdx.x;
    if (tid == 0) { sm.tile = atomicAdd(tile_counter, 1u); sm.status = tile_status; sm.in = keys; sm.n = n; }
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) sm.warp_hist[w][tid] = 0;
    if ((int)tid < nspl) sm.spl[tid] = __ldg(splitters + tid);
    const int nd = nspl + 1;
    sm.dst[tid] = ((int)tid < nd) ? __ldg(table + tid) : 0ull;
    __syncthreads();
    const SplitterDigit dg{sm.spl, nspl, steps};
    if ((long long)(sm.tile + 1) * SORT_TILE <= n)
        sweep_tile<SplitterDigit, true, true>(sm, sstride, nd, dg, steps);
    else
        sweep_tile<SplitterDigit, false, true>(sm, sstride, nd, dg, steps);
}

// The same exchange step for <= PB_MAX_PARTS destinations (one per GPU of a box) with BULK stores: after the tile has
// been ordered by destination in shared memory, each destination's run (~4096 / parts keys) leaves the SM as ONE
// cp.async.bulk shared -> global copy (TMA engine; 16-byte granules) instead of ~parts x 16 warp-wide 4-byte store
// instructions -- over NVLink that means few large writes per tile instead of thousands of 128-byte ones.  A bulk copy
// needs source and destination 16-byte aligned, so the look-back runs BEFORE the shared-memory scatter here (it tells
// where in the destination buffer the run starts) and every run is placed in shared memory at an offset congruent to its
// destination address modulo 16 bytes; the <= 3 keys before the first and after the last 16-byte granule are stored
// individually.  Ranking (ballot matching on the few destination bits) is the tile routine's.
constexpr int PB_MAX_PARTS = 32;
constexpr int PB_KEYS = SORT_TILE + 8 * PB_MAX_PARTS;          // room for the alignment gaps between runs (<= 6 keys each)

struct PartSmem {
    __align__(16) uint32_t keys[PB_KEYS];
    unsigned warp_hist[SORT_WARPS][PB_MAX_PARTS];     // per-warp destination counts -> scatter bases
    unsigned long long dst[PB_MAX_PARTS];             // byte address of this tile's run in destination d
    unsigned start[PB_MAX_PARTS], count[PB_MAX_PARTS];
    uint32_t spl[PB_MAX_PARTS];
    unsigned tile;
};

__device__ __forceinline__ void bulk_store_1d(unsigned long long gdst, uint32_t smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}

template <bool FULL>
__device__ __forceinline__ void partition_tile_bulk(PartSmem &sm, const uint32_t *__restrict__ keys_in, long long n,
                                                    const unsigned long long *__restrict__ table,
                                                    unsigned long long *status, int sstride, int nd, int steps) {
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tile = sm.tile;
    const int tile_n = FULL ? SORT_TILE : (int)(n - (long long)tile * SORT_TILE);
    const SplitterDigit digit_of{sm.spl, nd - 1, steps};

    uint32_t key[SORT_IPT];
    const int wbase = warp * (32 * SORT_IPT) + lane;
    const uint32_t *kp = keys_in + (long long)tile * SORT_TILE + wbase;
#pragma unroll
    for (int i = 0; i < SORT_IPT; i++) key[i] = (FULL || wbase + i * 32 < tile_n) ? __ldg(kp + i * 32) : 0u;
    unsigned dpk[SORT_IPT / 4];
#pragma unroll
    for (int i = 0; i < SORT_IPT; i++) {
        const unsigned d = digit_of(key[i]);
        dpk[i >> 2] = (i & 3) ? (dpk[i >> 2] | (d << (8 * (i & 3)))) : d;
    }
    auto dig = [&](int i) -> unsigned { return (dpk[i >> 2] >> (8 * (i & 3))) & 255u; };

    // rank inside the warp (stable): ballots on the destination bits, running per-destination counters
    unsigned rank2[SORT_IPT / 2];
    const unsigned lt = lanemask_lt();
    unsigned *wh = sm.warp_hist[warp];
#pragma unroll
    for (int i = 0; i < SORT_IPT; i++) {
        const bool valid = FULL || wbase + i * 32 < tile_n;
        const unsigned d = valid ? dig(i) : 0u;
        unsigned pm = __ballot_sync(0xffffffffu, valid);
        if (!valid) pm = ~pm;
        for (int b = 0; b < steps; b++) {
            const bool bit = (d >> b) & 1u;
            const unsigned mm = __ballot_sync(0xffffffffu, bit);
            pm &= bit ? mm : ~mm;
        }
        const unsigned below = __popc(pm & lt);
        unsigned old = 0;
        if (below == 0 && valid) {
            old = wh[d];
            wh[d] = old + __popc(pm);
        }
        old = __shfl_sync(0xffffffffu, old, __ffs(pm) - 1);
        const unsigned r16 = old + below;
        rank2[i >> 1] = (i & 1) ? (rank2[i >> 1] | (r16 << 16)) : r16;
        __syncwarp();
    }
    __syncthreads();

    // thread d < nd: tile count, publish, look back, destination address of this tile's run
    if ((int)tid < nd) {
        const unsigned d = tid;
        unsigned tot = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) tot += sm.warp_hist[w][d];
        unsigned long long *my = status + (size_t)tile * sstride + d;
        st_status(my, (tile == 0 ? FLAG_INC : FLAG_AGG) | tot);
        unsigned long long prefix = 0;
        if (tile > 0) {
            int t = (int)tile - 1;
            const unsigned long long *p = my - sstride;
            for (;;) {
                unsigned long long stt[LB_WINDOW];
#pragma unroll
                for (int j = 0; j < LB_WINDOW; j++) stt[j] = (j <= t) ? ld_status(p - (size_t)j * sstride) : FLAG_INC;
                unsigned lowest = 0xffffffffu;
#pragma unroll
                for (int j = 0; j < LB_WINDOW; j++) lowest = min(lowest, (unsigned)(stt[j] >> 32));
                if (lowest < 0x40000000u) continue;                 // some predecessor has not even counted yet: read again
                bool done = false;
#pragma unroll
                for (int j = 0; j < LB_WINDOW; j++) {
                    if (!done) prefix += stt[j] & VAL_MASK;
                    done = done || (unsigned)(stt[j] >> 32) >= 0x80000000u;
                }
                if (done) break;
                t -= LB_WINDOW;
                p -= (size_t)LB_WINDOW * sstride;
            }
            st_status(my, FLAG_INC | (prefix + tot));
        }
        sm.dst[d] = __ldg(table + d) + 4ull * prefix;
        sm.count[d] = tot;
    }
    __syncthreads();
    // shared-memory layout: run d starts at a 4-key boundary plus (destination address / 4) mod 4
    if (tid == 0) {
        unsigned cur = 0;
        for (int d = 0; d < nd; d++) {
            const unsigned st0 = ((cur + 3u) & ~3u) + (unsigned)((sm.dst[d] >> 2) & 3ull);
            sm.start[d] = st0;
            cur = st0 + sm.count[d];
        }
    }
    __syncthreads();
    if ((int)tid < nd) {                                   // scatter base of (warp w, destination d)
        unsigned run = sm.start[tid];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) { const unsigned c = sm.warp_hist[w][tid]; sm.warp_hist[w][tid] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_IPT; i++) {
        if (FULL || wbase + i * 32 < tile_n) {
            const unsigned r16 = (i & 1) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xffffu);
            sm.keys[wh[dig(i)] + r16] = key[i];
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> visible to the bulk-copy engine
    __syncthreads();
    // one thread per destination: <= 3 head keys, one bulk copy, <= 3 tail keys
    if ((int)tid < nd) {
        // (a zero table entry = the destination refused the reservation: its run is dropped, the owner reports the overflow)
        const unsigned cnt = __ldg(table + tid) ? sm.count[tid] : 0u, st0 = sm.start[tid];
        const unsigned long long a = sm.dst[tid];
        const unsigned head = min(cnt, (unsigned)(((16ull - (a & 15ull)) & 15ull) >> 2));
        const unsigned body = (cnt - head) & ~3u;
        for (unsigned j = 0; j < head; j++) *reinterpret_cast<uint32_t *>(a + 4ull * j) = sm.keys[st0 + j];
        if (body) bulk_store_1d(a + 4ull * head, smem_u32(&sm.keys[st0 + head]), body * 4u);
        for (unsigned j = head + body; j < cnt; j++) *reinterpret_cast<uint32_t *>(a + 4ull * j) = sm.keys[st0 + j];
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // shared memory must outlive the copy's reads
    }
}

__global__ void __launch_bounds__(SORT_THREADS, 4)
partition_scatter_bulk_kernel(const uint32_t *__restrict__ keys, long long n, const uint32_t *__restrict__ splitters, int nspl,
                              int steps, const unsigned long long *__restrict__ table, unsigned long long *tile_status,
                              int sstride, unsigned *tile_counter) {
    __shared__ PartSmem sm;
    const unsigned tid = threadIdx.x;
    if (tid == 0) sm.tile = atomicAdd(tile_counter, 1u);
    if (tid < SORT_WARPS * PB_MAX_PARTS) (&sm.warp_hist[0][0])[tid] = 0;
    if ((int)tid < nspl) sm.spl[tid] = __ldg(splitters + tid);
    __syncthreads();
    if ((long long)(sm.tile + 1) * SORT_TILE <= n)
        partition_tile_bulk<true>(sm, keys, n, table, tile_status, sstride, nspl + 1, steps);
    else
        partition_tile_bulk<false>(sm, keys, n, table, tile_status, sstride, nspl + 1, steps);
}

// ---- exchange by remote append ------------------------------------------------------------------------------
// The receive side of the exchange is itself an evaluator buffer (two key streams + state) in peer-mapped memory.  A tile
// orders its keys by destination in shared memory, RESERVES its run in every destination with one system-scope
// atomicAdd on that destination's stream counter (over NVLink for a peer), and stores the run with one bulk copy.
// Nothing has to be known in advance: no counting pass, no all-gather of bucket sizes, no look-back between tiles, and
// both streams go in one launch.  The order inside a destination stream is the order of the reservations, which is fine:
// a stream is a multiset.  Overflow: each stream is kept inside the destination buffer here; the two streams meeting in
// the middle is detected by the owner when it reads its state (as for local appends).
struct ExchangeDst {
    unsigned long long keys[PB_MAX_PARTS];     // byte address of destination d's key buffer
    unsigned long long state[PB_MAX_PARTS];    // byte address of destination d's EvalState
    long long capacity;                        // keys per destination buffer
};

// Per-phase cycle counters of exchange_tile (development builds only: MSS_NVCC_EXTRA=-DMSS_EXCH_PROFILE; tools/exch_phases.py)
#ifdef MSS_EXCH_PROFILE
__device__ unsigned long long g_exch_prof[32];
__device__ __forceinline__ unsigned long long prof_clock() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}
#define EXCH_T(var) const unsigned long long var = prof_clock()
#define EXCH_ADD(i, d) atomicAdd(&g_exch_prof[i], (unsigned long long)(d))
extern "C" MSS_API int mss_debug_exchange_profile(unsigned long long *out32_host, int reset) {
    if (out32_host) cudaMemcpyFromSymbol(out32_host, g_exch_prof, sizeof(g_exch_prof));
    if (reset) { unsigned long long z[32] = {}; cudaMemcpyToSymbol(g_exch_prof, z, sizeof(z)); }
    return 0;
}
#else
#define EXCH_T(var)
#define EXCH_ADD(i, d)
#endif

template <bool FULL, bool SMALL>
__device__ __forceinline__ void exchange_tile(PartSmem &sm, const uint32_t *__restrict__ keys_in, long long n, unsigned tile,
                                              bool positives, const ExchangeDst &dst, int nd, int steps) {
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_n = FULL ? SORT_TILE : (int)(n - (long long)tile * SORT_TILE);
    const SplitterDigit digit_of{sm.spl, nd - 1, steps};
    EXCH_T(t0);

    uint32_t key[SORT_IPT];
    const int wbase = warp * (32 * SORT_IPT) + lane;
    const uint32_t *kp = keys_in + (long long)tile * SORT_TILE + wbase;
#pragma unroll
    for (int i = 0; i < SORT_IPT; i++) key[i] = (FULL || wbase + i * 32 < tile_n) ? __ldg(kp + i * 32) : 0u;
    unsigned dpk[SORT_IPT / 4];
    unsigned *wh = sm.warp_hist[warp];
    unsigned rank2[SORT_IPT / 2];
    if (SMALL) {
        // <= 8 destinations (round 2: the ballot ranking below costs 3.4 warp instructions per key, which made this kernel
        // issue-bound at 250 Gkeys/s).  Count per thread instead -- 8-bit fields of a 64-bit word, a thread has 16 keys --
        // widen to 16-bit fields, ONE packed warp scan (4 words), and a key's rank is its thread's exclusive offset for
        // that destination plus a running count.  The order inside a destination becomes thread-major, which is as good
        // as any: a stream is a multiset.
        uint32_t sp[7];
#pragma unroll
        for (int j = 0; j < 7; j++) sp[j] = j < nd - 1 ? sm.spl[j] : 0xFFFFFFFFu;
        unsigned long long c8 = 0;
#pragma unroll
        for (int i = 0; i < SORT_IPT; i++) {
            unsigned d = 0;
#pragma unroll
            for (int j = 0; j < 7; j++) d += key[i] >= sp[j] ? 1u : 0u;
            d = min(d, (unsigned)(nd - 1));                // (a key of 0xFFFFFFFF would pass the padding: a NaN, flagged anyway)
            dpk[i >> 2] = (i & 3) ? (dpk[i >> 2] | (d << (8 * (i & 3)))) : d;
            if (FULL || wbase + i * 32 < tile_n) c8 += 1ull << (8 * d);
        }
        // 16-bit fields: e[0] = (d0, d2), e[1] = (d1, d3), e[2] = (d4, d6), e[3] = (d5, d7)
        const unsigned lo = (unsigned)c8, hi = (unsigned)(c8 >> 32);
        unsigned c[4] = {lo & 0x00FF00FFu, (lo >> 8) & 0x00FF00FFu, hi & 0x00FF00FFu, (hi >> 8) & 0x00FF00FFu};
        unsigned e[4] = {c[0], c[1], c[2], c[3]};
#pragma unroll
        for (int s2 = 1; s2 < 32; s2 <<= 1) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned u = __shfl_up_sync(0xffffffffu, e[q], s2);
                if (lane >= s2) e[q] += u;
            }
        }
        if (lane == 31) {                                  // the warp's counts per destination
            wh[0] = e[0] & 0xFFFFu; wh[2] = e[0] >> 16; wh[1] = e[1] & 0xFFFFu; wh[3] = e[1] >> 16;
            wh[4] = e[2] & 0xFFFFu; wh[6] = e[2] >> 16; wh[5] = e[3] & 0xFFFFu; wh[7] = e[3] >> 16;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) e[q] -= c[q];          // exclusive: keys of the lower lanes, per destination
#pragma unroll
        for (int i = 0; i < SORT_IPT; i++) {
            const unsigned d = (dpk[i >> 2] >> (8 * (i & 3))) & 255u;
            const unsigned w = ((d >> 1) & 2u) | (d & 1u), sh = (d & 2u) << 3;      // word e[w], field at bit sh
            const unsigned x = w == 0 ? e[0] : w == 1 ? e[1] : w == 2 ? e[2] : e[3];
            const unsigned r16 = (x >> sh) & 0xFFFFu;
            const unsigned inc = (FULL || wbase + i * 32 < tile_n) ? (1u << sh) : 0u;
            e[0] += w == 0 ? inc : 0u; e[1] += w == 1 ? inc : 0u; e[2] += w == 2 ? inc : 0u; e[3] += w == 3 ? inc : 0u;
            rank2[i >> 1] = (i & 1) ? (rank2[i >> 1] | (r16 << 16)) : r16;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SORT_IPT; i++) {
            const unsigned d = digit_of(key[i]);
            dpk[i >> 2] = (i & 3) ? (dpk[i >> 2] | (d << (8 * (i & 3)))) : d;
        }
    }
    auto dig = [&](int i) -> unsigned { return (dpk[i >> 2] >> (8 * (i & 3))) & 255u; };
    EXCH_T(t1);                                            // keys loaded (the digits depend on them)

    if (!SMALL) {
        const unsigned lt = lanemask_lt();
#pragma unroll
        for (int i = 0; i < SORT_IPT; i++) {
            const bool valid = FULL || wbase + i * 32 < tile_n;
            const unsigned d = valid ? dig(i) : 0u;
            unsigned pm = __ballot_sync(0xffffffffu, valid);
            if (!valid) pm = ~pm;
            for (int b = 0; b < steps; b++) {
                const bool bit = (d >> b) & 1u;
                const unsigned mm = __ballot_sync(0xffffffffu, bit);
                pm &= bit ? mm : ~mm;
            }
            const unsigned below = __popc(pm & lt);
            unsigned old = 0;
            if (below == 0 && valid) {
                old = wh[d];
                wh[d] = old + __popc(pm);
            }
            old = __shfl_sync(0xffffffffu, old, __ffs(pm) - 1);
            const unsigned r16 = old + below;
            rank2[i >> 1] = (i & 1) ? (rank2[i >> 1] | (r16 << 16)) : r16;
            __syncwarp();
        }
    }
    EXCH_T(t2);                                            // ranked
    __syncthreads();
    EXCH_T(t3);

    // thread d < nd: reserve this tile's run in destination d (one remote atomic), address of the run
    if ((int)tid < nd) {
        const unsigned d = tid;
        unsigned tot = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) tot += sm.warp_hist[w][d];
        unsigned long long addr = 0;
        if (tot) {
            EvalState *es = reinterpret_cast<EvalState *>(dst.state[d]);
            const unsigned long long base = atomicAdd_system(positives ? &es->n_pos : &es->n_neg, (unsigned long long)tot);
            if (base + tot > (unsigned long long)dst.capacity) {
                atomicAdd_system(&es->overflow, (unsigned long long)tot);      // dropped: the owner reports MSS_ERR_WORKSPACE
                tot = 0;
            } else {
                // negatives fill the buffer upwards, positives downwards (the run itself is stored in ascending order)
                addr = dst.keys[d] + 4ull * (positives ? (unsigned long long)dst.capacity - base - tot : base);
            }
        }
        sm.dst[d] = addr;
        sm.count[d] = tot;
#ifdef MSS_EXCH_PROFILE
        { EXCH_T(ta); EXCH_ADD(16 + d, ta - t3); }         // reservation round trip per destination
#endif
    }
    __syncthreads();
    EXCH_T(t4);
    if (tid == 0) {
        unsigned cur = 0;
        for (int d = 0; d < nd; d++) {
            unsigned c = 0;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; w++) c += sm.warp_hist[w][d];       // (a dropped run still needs its shared-memory room)
            const unsigned st0 = ((cur + 3u) & ~3u) + (unsigned)((sm.dst[d] >> 2) & 3ull);
            sm.start[d] = st0;
            cur = st0 + c;
        }
    }
    __syncthreads();
    if ((int)tid < nd) {
        unsigned run = sm.start[tid];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) { const unsigned c = sm.warp_hist[w][tid]; sm.warp_hist[w][tid] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_IPT; i++) {
        if (FULL || wbase + i * 32 < tile_n) {
            const unsigned r16 = (i & 1) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xffffu);
            sm.keys[wh[dig(i)] + r16] = key[i];
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    EXCH_T(t5);
    if ((int)tid < nd) {
        const unsigned cnt = sm.count[tid], st0 = sm.start[tid];
        const unsigned long long a = sm.dst[tid];
        const unsigned head = min(cnt, (unsigned)(((16ull - (a & 15ull)) & 15ull) >> 2));
        const unsigned body = (cnt - head) & ~3u;
        for (unsigned j = 0; j < head; j++) *reinterpret_cast<uint32_t *>(a + 4ull * j) = sm.keys[st0 + j];
        if (body) bulk_store_1d(a + 4ull * head, smem_u32(&sm.keys[st0 + head]), body * 4u);
        for (unsigned j = head + body; j < cnt; j++) *reinterpret_cast<uint32_t *>(a + 4ull * j) = sm.keys[st0 + j];
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        EXCH_T(t6);
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#ifdef MSS_EXCH_PROFILE
        { EXCH_T(t7); EXCH_ADD(24 + tid, t7 - t6); if (tid == 0) EXCH_ADD(6, t6 - t5); }   // bulk read-completion per destination
#endif
    }
#ifdef MSS_EXCH_PROFILE
    if (tid == 0) {
        EXCH_T(t8);
        EXCH_ADD(0, t1 - t0); EXCH_ADD(1, t2 - t1); EXCH_ADD(2, t3 - t2); EXCH_ADD(3, t4 - t3); EXCH_ADD(4, t5 - t4);
        EXCH_ADD(7, t8 - t0); EXCH_ADD(15, 1);
    }
#endif
}

template <bool FULL>
__device__ __forceinline__ void exchange_tile_any(PartSmem &sm, const uint32_t *__restrict__ keys_in, long long n, unsigned tile,
                                                  bool positives, const ExchangeDst &dst, int nd, int steps) {
    if (nd <= 8) exchange_tile<FULL, true>(sm, keys_in, n, tile, positives, dst, nd, steps);
    else exchange_tile<FULL, false>(sm, keys_in, n, tile, positives, dst, nd, steps);
}

// blockIdx.x < tiles_neg: a tile of the in-distribution stream, else of the OOD stream
__global__ void __launch_bounds__(SORT_THREADS, 4)
exchange_append_kernel(const uint32_t *__restrict__ neg, long long n_neg, const uint32_t *__restrict__ pos, long long n_pos,
                       unsigned tiles_neg, const uint32_t *__restrict__ splitters, int nspl, int steps, ExchangeDst dst) {
    __shared__ PartSmem sm;
    const unsigned tid = threadIdx.x;
    if (tid < SORT_WARPS * PB_MAX_PARTS) (&sm.warp_hist[0][0])[tid] = 0;
    if ((int)tid < nspl) sm.spl[tid] = __ldg(splitters + tid);
    __syncthreads();
    const bool positives = blockIdx.x >= tiles_neg;
    const unsigned tile = positives ? blockIdx.x - tiles_neg : blockIdx.x;
    const uint32_t *keys = positives ? pos : neg;
    const long long n = positives ? n_pos : n_neg;
    if ((long long)(tile + 1) * SORT_TILE <= n) exchange_tile_any<true>(sm, keys, n, tile, positives, dst, nspl + 1, steps);
    else exchange_tile_any<false>(sm, keys, n, tile, positives, dst, nspl + 1, steps);
}

// The same, with the stream sizes read from the source evaluator's DEVICE state: nothing on the host has to know how
// many keys a batch produced, so a batch can be exchanged right behind the kernel that scored it, on a side stream, while
// the next batch is being scored (StreamingEvaluator(exchange="stream")).  gridDim.x bounds the number of tiles.
__global__ void __launch_bounds__(SORT_THREADS, 4)
exchange_append_dev_kernel(const uint32_t *__restrict__ keys, long long capacity, const EvalState *__restrict__ state,
                           const uint32_t *__restrict__ splitters, int nspl, int steps, ExchangeDst dst) {
    __shared__ PartSmem sm;
    const unsigned tid = threadIdx.x;
    const long long n_neg = (long long)min(state->n_neg, (unsigned long long)capacity);
    const long long n_pos = (long long)min(state->n_pos, (unsigned long long)(capacity - n_neg));
    const unsigned tiles_neg = (unsigned)((n_neg + SORT_TILE - 1) / SORT_TILE), tiles_pos = (unsigned)((n_pos + SORT_TILE - 1) / SORT_TILE);
    if ((int)tid < nspl) sm.spl[tid] = __ldg(splitters + tid);
    // grid-stride over the tiles: the host caps the grid (a few CTAs per SM) so that this kernel, which mostly waits for
    // remote reservations, only ever holds a fraction of an SM's slots next to the scoring kernel it runs beside
    for (unsigned t = blockIdx.x; t < tiles_neg + tiles_pos; t += gridDim.x) {
        if (tid < SORT_WARPS * PB_MAX_PARTS) (&sm.warp_hist[0][0])[tid] = 0;
        __syncthreads();
        const bool positives = t >= tiles_neg;
        const unsigned tile = positives ? t - tiles_neg : t;
        const uint32_t *src = positives ? keys + (capacity - n_pos) : keys;
        const long long n = positives ? n_pos : n_neg;
        if ((long long)(tile + 1) * SORT_TILE <= n) exchange_tile_any<true>(sm, src, n, tile, positives, dst, nspl + 1, steps);
        else exchange_tile_any<false>(sm, src, n, tile, positives, dst, nspl + 1, steps);
        __syncthreads();                                   // the bulk copies have read the tile's shared memory (wait_group.read)
    }
}

// accum += staging (counts, flags); staging = 0 -- the staging buffer is empty again for the next batch
__global__ void eval_state_fold_kernel(EvalState *accum, EvalState *staging) {
    accum->n_neg += staging->n_neg;
    accum->n_pos += staging->n_pos;
    accum->nan_flag |= staging->nan_flag;
    accum->inf_flag |= staging->inf_flag;
    accum->overflow += staging->overflow;
    staging->n_neg = 0; staging->n_pos = 0; staging->nan_flag = 0; staging->inf_flag = 0; staging->overflow = 0;
}

// ---- streamed exchange through the copy engines ---------------------------------------------------------------
// Measured on B200 (round 2, tools/probe_stream.py, 2 GPUs, 1.0 G keys per rank): the remote-append kernel beside the
// scoring kernel costs the scoring loop +12 ms, the same kernel with LOCAL destinations +9 ms -- but it runs in 2.9 ms
// when it has the GPU to itself (two issue-bound kernels sharing the SMs lose more than they overlap), and 8.6 ms alone
// with remote destinations (NVLink at 230 GB/s: small runs, one round trip per tile).  So the staged form: partition a
// batch into LOCAL per-destination outboxes right behind its scoring kernel, on the same stream; one thread per
// (destination, stream) then reserves the run in the owner's receive buffer with a single system-scope atomicAdd and
// writes (count, offset) to pinned host memory; the host hands the 2 x ranks block copies of the batch to the COPY
// ENGINES (mss_memcpy_async), which move them over NVLink while the SMs score the next batch.
__global__ void exchange_plan_kernel(ExchangeDst outbox, ExchangeDst recv, int parts, unsigned long long *plan) {
    const int t = threadIdx.x;
    if (t >= 2 * parts) return;
    const int d = t >> 1, positives = t & 1;
    EvalState *os = reinterpret_cast<EvalState *>(outbox.state[d]);
    EvalState *rs = reinterpret_cast<EvalState *>(recv.state[d]);
    unsigned long long cnt = positives ? os->n_pos : os->n_neg, off = 0;
    if (!positives && os->overflow) atomicAdd_system(&rs->overflow, os->overflow);     // an outbox overflowed: the owner reports it
    if (cnt) {
        const unsigned long long base = atomicAdd_system(positives ? &rs->n_pos : &rs->n_neg, cnt);
        if (base + cnt > (unsigned long long)recv.capacity) {
            atomicAdd_system(&rs->overflow, cnt);                                       // dropped: the owner reports MSS_ERR_WORKSPACE
            cnt = 0;
        } else {
            off = positives ? (unsigned long long)recv.capacity - base - cnt : base;    // first key of the run in the owner's buffer
        }
    }
    plan[4 * d + 2 * positives] = cnt;
    plan[4 * d + 2 * positives + 1] = off;
    __threadfence_system();
}

// ---- streamed exchange without per-tile remote atomics ------------------------------------------------------
// One batch (a staging evaluator whose sizes live in its device state): count its keys per destination, reserve ONE run
// per destination and stream with a system-scope atomicAdd (2 x ranks remote atomics per batch instead of one per tile
// and destination), then the look-back / bulk-store scatter at the reserved addresses.  No host involvement.
struct StagingView {
    const uint32_t *keys;
    long long capacity;
    const EvalState *state;
    __device__ __forceinline__ void stream(int positives, const uint32_t *&p, long long &n) const {
        const long long n_neg = (long long)min(state->n_neg, (unsigned long long)capacity);
        const long long n_pos = (long long)min(state->n_pos, (unsigned long long)(capacity - n_neg));
        p = positives ? keys + (capacity - n_pos) : keys;
        n = positives ? n_pos : n_neg;
    }
};

// blockIdx.y = stream; counts[stream][RADIX]; parts <= 16 (lane-private counters)
__global__ void __launch_bounds__(256)
partition_count_dev_kernel(StagingView sv, const uint32_t *__restrict__ splitters, int nspl, int steps,
                           unsigned long long *__restrict__ counts_all) {
    __shared__ unsigned s_c[PB_MAX_PARTS * 256];
    __shared__ uint32_t s_spl[PB_MAX_PARTS];
    const uint32_t *keys;
    long long n;
    sv.stream(blockIdx.y, keys, n);
    unsigned long long *counts = counts_all + blockIdx.y * RADIX;
    for (int i = threadIdx.x; i < PB_MAX_PARTS * 256; i += 256) s_c[i] = 0;
    if ((int)threadIdx.x < nspl) s_spl[threadIdx.x] = __ldg(splitters + threadIdx.x);
    __syncthreads();
    const SplitterDigit dg{s_spl, nspl, steps};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s_c[dg(__ldg(keys + i)) * 256 + threadIdx.x]++;
    __syncthreads();
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int p = warp; p <= nspl; p += 8) {
        unsigned long long acc = 0;
        for (int t = lane; t < 256; t += 32) acc += s_c[p * 256 + t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0 && acc) atomicAdd(counts + p, acc);
    }
}

// thread (stream, destination): reserve the batch's run, write its byte address (0 = refused) into table[stream][d]
__global__ void exchange_reserve_kernel(const unsigned long long *__restrict__ counts_all, ExchangeDst dst, int parts,
                                        unsigned long long *__restrict__ table_all) {
    const int sid = threadIdx.x / PB_MAX_PARTS, d = threadIdx.x % PB_MAX_PARTS;
    if (sid > 1 || d >= parts) return;
    const unsigned long long tot = counts_all[sid * RADIX + d];
    unsigned long long addr = dst.keys[d];                 // (an empty run still needs a non-zero entry)
    if (tot) {
        EvalState *es = reinterpret_cast<EvalState *>(dst.state[d]);
        const unsigned long long base = atomicAdd_system(sid ? &es->n_pos : &es->n_neg, tot);
        if (base + tot > (unsigned long long)dst.capacity) {
            atomicAdd_system(&es->overflow, tot);
            addr = 0;
        } else {
            addr = dst.keys[d] + 4ull * (sid ? (unsigned long long)dst.capacity - base - tot : base);
        }
    }
    table_all[sid * PB_MAX_PARTS + d] = addr;
}

// blockIdx.y = stream; tickets, status words and the address table per stream
__global__ void __launch_bounds__(SORT_THREADS, 4)
partition_scatter_bulk_dev_kernel(StagingView sv, const uint32_t *__restrict__ splitters, int nspl, int steps,
                                  const unsigned long long *__restrict__ table_all, unsigned long long *status_all,
                                  size_t status_words, int sstride, unsigned *counters) {
    __shared__ PartSmem sm;
    const unsigned tid = threadIdx.x;
    const uint32_t *keys;
    long long n;
    sv.stream(blockIdx.y, keys, n);
    if (tid == 0) sm.tile = atomicAdd(counters + blockIdx.y, 1u);
    if (tid < SORT_WARPS * PB_MAX_PARTS) (&sm.warp_hist[0][0])[tid] = 0;
    if ((int)tid < nspl) sm.spl[tid] = __ldg(splitters + tid);
    __syncthreads();
    if ((long long)sm.tile * SORT_TILE >= n) return;       // tickets beyond the batch (the grid is sized for the capacity)
    const unsigned long long *table = table_all + blockIdx.y * PB_MAX_PARTS;
    unsigned long long *status = status_all + blockIdx.y * status_words;
    if ((long long)(sm.tile + 1) * SORT_TILE <= n)
        partition_tile_bulk<true>(sm, keys, n, table, status, sstride, nspl + 1, steps);
    else
        partition_tile_bulk<false>(sm, keys, n, table, status, sstride, nspl + 1, steps);
}

// top-`bits` histogram for splitter selection (bins = 1 << bits <= 65536).
// First form: warp-aggregated atomics straight to the global bins -- 465 ms for 2 G keys on B200 (cfg-4, two
// ranks): real score distributions put most keys into a few hundred bins, and every SM hammered the same L2
// addresses.  Now each CTA (one per SM) privatises a window of 32 768 bins in 128 KB of shared memory, walks its
// contiguous chunk of keys once per window (two windows for 16 bits), and adds its non-zero bins to the global
// histogram at the end of the window.  Inside a warp, groups of equal bins are peeled with ballots so that a
// hot bin costs one shared-memory atomic per warp, not 32 serialised ones.
constexpr int KH_THREADS = 512;
constexpr int KH_WINDOW = 32768;
constexpr int KH_SMEM = KH_WINDOW * 4;

__device__ __forceinline__ void kh_add(unsigned *s_hist, unsigned b, bool valid) {
    unsigned remaining = __ballot_sync(0xffffffffu, valid);
    const unsigned lane = lane_id();
#pragma unroll 1
    for (int round = 0; round < 4 && remaining; round++) {
        const int leader = __ffs(remaining) - 1;
        const unsigned v = __shfl_sync(0xffffffffu, b, leader);
        const unsigned m = __ballot_sync(0xffffffffu, valid && b == v) & remaining;
        if ((int)lane == leader) atomicAdd(s_hist + v, (unsigned)__popc(m));
        remaining &= ~m;
    }
    if ((remaining >> lane) & 1u) atomicAdd(s_hist + b, 1u);
}

__global__ void __launch_bounds__(KH_THREADS, 1)
keys_histogram_kernel(const uint32_t *__restrict__ keys, long long n, int shift, unsigned nbins, int every,
                      unsigned long long *__restrict__ hist, int refine, unsigned prefix) {
    // refine: second-level histogram -- only keys whose top 16 bits equal `prefix` are counted, by their LOW 16 bits
    // (a splitter that has to fall inside a heavy coarse bin is placed with this resolution)
    extern __shared__ __align__(16) unsigned s_hist[];
    __shared__ unsigned s_outside;      // keys of this CTA's chunk that fell outside the current window
    // contiguous chunk of whole uint4 groups per CTA; the <= 3 + 3 unaligned head / tail keys go to CTA 0
    const long long head = min(n, (long long)(((16 - ((uintptr_t)keys & 15)) & 15) >> 2));
    const long long groups_all = (n - head) >> 2;
    // every > 1: systematic sample -- only uint4 group number j * every (j = 0, 1, ...) is counted
    const long long groups4 = (groups_all + every - 1) / every;
    const uint4 *k4 = reinterpret_cast<const uint4 *>(keys + head);
    const long long per = (groups4 + gridDim.x - 1) / gridDim.x;
    const long long g0 = min(groups4, (long long)blockIdx.x * per), g1 = min(groups4, g0 + per);
    const long long iters = (g1 - g0 + KH_THREADS - 1) / KH_THREADS;          // uniform per CTA (ballots inside)
    bool more = true;                                                          // CTA-uniform
    for (unsigned base = 0; base < nbins && more; base += KH_WINDOW) {
        for (int i = threadIdx.x; i < KH_WINDOW; i += KH_THREADS) s_hist[i] = 0;
        if (threadIdx.x == 0) s_outside = 0;
        __syncthreads();
        unsigned outside = 0;
        for (long long it = 0; it < iters; it++) {
            const long long g = g0 + it * KH_THREADS + threadIdx.x;
            const bool in = g < g1;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (in) v = __ldg(k4 + g * every);
            unsigned b0, b1, b2, b3;
            if (refine) {
                // keys outside the prefix get a bin that is in no window and is not counted as "outside" either
                const bool p0 = (v.x >> 16) == prefix, p1 = (v.y >> 16) == prefix, p2 = (v.z >> 16) == prefix,
                           p3 = (v.w >> 16) == prefix;
                b0 = p0 ? (v.x & 0xffffu) - base : 0xffffffffu; b1 = p1 ? (v.y & 0xffffu) - base : 0xffffffffu;
                b2 = p2 ? (v.z & 0xffffu) - base : 0xffffffffu; b3 = p3 ? (v.w & 0xffffu) - base : 0xffffffffu;
                if (in) outside += (p0 && b0 >= (unsigned)KH_WINDOW) + (p1 && b1 >= (unsigned)KH_WINDOW) +
                                   (p2 && b2 >= (unsigned)KH_WINDOW) + (p3 && b3 >= (unsigned)KH_WINDOW);
            } else {
                b0 = (v.x >> shift) - base; b1 = (v.y >> shift) - base; b2 = (v.z >> shift) - base; b3 = (v.w >> shift) - base;
                // the lane's own four keys first: adjacent pixels usually share a bin
                if (in) outside += (b0 >= (unsigned)KH_WINDOW) + (b1 >= (unsigned)KH_WINDOW) + (b2 >= (unsigned)KH_WINDOW) +
                                   (b3 >= (unsigned)KH_WINDOW);
            }
            const bool same4 = b0 == b1 && b0 == b2 && b0 == b3;
            if (__all_sync(0xffffffffu, same4 || !in)) {
                // every lane holds four equal bins: one aggregation round with weight 4
                const bool valid = in && b0 < (unsigned)KH_WINDOW;
                unsigned remaining = __ballot_sync(0xffffffffu, valid);
                const unsigned lane = lane_id();
#pragma unroll 1
                for (int round = 0; round < 4 && remaining; round++) {
                    const int leader = __ffs(remaining) - 1;
                    const unsigned vv = __shfl_sync(0xffffffffu, b0, leader);
                    const unsigned m = __ballot_sync(0xffffffffu, valid && b0 == vv) & remaining;
                    if ((int)lane == leader) atomicAdd(s_hist + vv, 4u * (unsigned)__popc(m));
                    remaining &= ~m;
                }
                if ((remaining >> lane) & 1u) atomicAdd(s_hist + b0, 4u);
            } else {
                kh_add(s_hist, b0, in && b0 < (unsigned)KH_WINDOW);
                kh_add(s_hist, b1, in && b1 < (unsigned)KH_WINDOW);
                kh_add(s_hist, b2, in && b2 < (unsigned)KH_WINDOW);
                kh_add(s_hist, b3, in && b3 < (unsigned)KH_WINDOW);
            }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && every == 1)
            for (long long i = 0; i < n; i++) {
                if (i == head) i += groups_all << 2;
                if (i >= n) break;
                const uint32_t kk = __ldg(keys + i);
                if (refine && (kk >> 16) != prefix) continue;
                const unsigned b = (refine ? (kk & 0xffffu) : (kk >> shift)) - base;
                if (b < (unsigned)KH_WINDOW) atomicAdd(s_hist + b, 1u);
                else outside++;
            }
        if (outside) atomicAdd(&s_outside, 1u);
        __syncthreads();
        // windows are visited in ascending order and the unsigned compare also counts the bins BELOW the window
        // as outside, so "nothing outside" means every key of the chunk has been counted: skip the rest
        more = s_outside != 0;
        for (int i = threadIdx.x; i < KH_WINDOW && base + i < nbins; i += KH_THREADS)
            if (s_hist[i]) atomicAdd(hist + base + i, (unsigned long long)s_hist[i]);
        __syncthreads();
    }
}

// test hooks: MSS_HIST_EPOCH=<chunks> shortens the counter-fold period so small inputs exercise it;
// MSS_SORT_NOSKIP=1 disables the single-bin pass skip (A/B timing, parity cross-check)
static int hist_epoch_chunks() {
    static const int v = [] {
        const char *e = getenv("MSS_HIST_EPOCH");
        const int x = e ? atoi(e) : HIST_EPOCH;
        return (x >= 1 && x <= HIST_EPOCH) ? x : HIST_EPOCH;
    }();
    return v;
}
static int sort_allow_skip() {
    static const int v = [] {
        const char *e = getenv("MSS_SORT_NOSKIP");
        return (e && atoi(e)) ? 0 : 1;
    }();
    return v;
}

static size_t sort_tiles(int64_t n) { return (size_t)((n + SORT_TILE - 1) / SORT_TILE); }

// kernels that opt into > 48 KB of dynamic shared memory: the attribute is per DEVICE, so it is (cheaply) set before
// every launch instead of once per process (a process may touch a second GPU)
static cudaError_t opt_in_smem() {
    cudaError_t e = cudaFuncSetAttribute(radix_histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HIST_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(keys_histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KH_SMEM);
}

struct SortWs {
    uint32_t *alt;                 // [n_upper + 8]
    SortPlan *plan;
    unsigned long long *hist;      // [2][4][256]
    unsigned *counters;            // [4] (+pad)
    unsigned long long *status;    // [4][tiles_upper][256]
    size_t zero_bytes;             // hist..status are contiguous: one memset
    char *zero_base;
    size_t tiles_upper;
};

static bool carve_sort(void *ws, size_t bytes, int64_t n_upper, SortWs &o) {
    Carver c(ws, bytes);
    o.tiles_upper = sort_tiles(n_upper) + 2;           // two segments: up to one partial tile each
    o.alt = c.take<uint32_t>((size_t)n_upper + 8);
    o.plan = c.take<SortPlan>(1);
    o.hist = c.take<unsigned long long>(2 * 4 * RADIX);
    o.zero_base = (char *)o.hist;
    o.counters = c.take<unsigned>(64);
    o.status = c.take<unsigned long long>(4 * o.tiles_upper * RADIX);
    o.zero_bytes = (size_t)((char *)(o.status + 4 * o.tiles_upper * RADIX) - o.zero_base);
    return c.ok();
}

size_t sort_ws_bytes(int64_t n_upper) {
    if (n_upper < 0) n_upper = 0;
    return align_up(((size_t)n_upper + 8) * 4, 256) + 256 + 2 * 4 * RADIX * 8 + 256 + 256 +
           4 * (sort_tiles(n_upper) + 2) * RADIX * 8 + 2048;
}

int sort_enqueue(const mss_eval_buffers *ev, uint32_t *xa, int64_t na, uint32_t *xb, int64_t nb, int64_t n_upper,
                 void *ws, size_t ws_bytes, cudaStream_t st, const SortPlan **plan_dev) {
    SortWs w;
    if (!carve_sort(ws, ws_bytes, n_upper, w)) {
        set_error("sort: workspace too small (%zu < %zu)", ws_bytes, sort_ws_bytes(n_upper));
        return MSS_ERR_WORKSPACE;
    }
    MSS_REQUIRE(w.tiles_upper < (1ull << 31), "sort: n too large");
    // histograms + tile counters by a small memset; the (large) status arrays are zeroed by the histogram kernel
    const size_t small_zero = (size_t)((char *)w.status - w.zero_base);
    MSS_CHECK_CUDA(cudaMemsetAsync(w.zero_base, 0, n_upper > 0 ? small_zero : w.zero_bytes, st));
    if (ev)
        sort_plan_kernel<<<1, 1, 0, st>>>(w.plan, nullptr, nullptr, 0, nullptr, nullptr, 0, (const EvalState *)ev->state,
                                          ev->keys, w.alt, ev->capacity);
    else
        sort_plan_kernel<<<1, 1, 0, st>>>(w.plan, xa, w.alt, na, xb, w.alt + ((na + 3) & ~3ll), nb, nullptr, nullptr,
                                          nullptr, 0);
    MSS_CHECK_LAUNCH();
    if (plan_dev) *plan_dev = w.plan;
    if (n_upper <= 0) return MSS_OK;
    MSS_CHECK_CUDA(opt_in_smem());
    // one CTA per SM and segment (192 KB of private counters each); a segment uses only the CTAs it has chunks for
    const int hgrid = (int)std::max<long long>(1, std::min<long long>((n_upper + 4 * HIST_CHUNK_KEYS - 1) / (4 * HIST_CHUNK_KEYS),
                                                                      (long long)sm_count()));
    radix_histogram_kernel<<<dim3(hgrid, 2), HIST_THREADS, HIST_SMEM, st>>>(w.plan, w.hist, hist_epoch_chunks(), (uint4 *)w.status,
                                                                            (w.zero_bytes - small_zero) / 16);
    MSS_CHECK_LAUNCH();
    radix_scan_bins_kernel<<<1, RADIX, 0, st>>>(w.hist, w.plan, sort_allow_skip());
    MSS_CHECK_LAUNCH();
    for (int p = 0; p < 4; p++) {
        onesweep_pass_kernel<<<(unsigned)w.tiles_upper, SORT_THREADS, 0, st>>>(
            w.plan, p, w.hist, w.status + (size_t)p * w.tiles_upper * RADIX, w.counters + p);
        MSS_CHECK_LAUNCH();
    }
    const int cgrid = (int)std::min<long long>((n_upper + 1023) / 1024, (long long)sm_count() * 8);
    sort_copy_back_kernel<<<std::max(cgrid, 1), 256, 0, st>>>(w.plan);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

int plan_enqueue(SortPlan *plan_dev, const uint32_t *xa, int64_t na, const uint32_t *xb, int64_t nb, cudaStream_t st) {
    sort_plan_kernel<<<1, 1, 0, st>>>(plan_dev, const_cast<uint32_t *>(xa), nullptr, na, const_cast<uint32_t *>(xb), nullptr, nb,
                                      nullptr, nullptr, nullptr, 0);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

}  // namespace mss

using namespace mss;

extern "C" size_t mss_sort_keys_workspace_bytes(int64_t n_total) { return sort_ws_bytes(n_total); }

extern "C" int mss_sort_keys(uint32_t *keys_a, int64_t n_a, uint32_t *keys_b, int64_t n_b, void *workspace,
                             size_t workspace_bytes, void *stream) {
    MSS_REQUIRE(n_a >= 0 && n_b >= 0, "mss_sort_keys: negative size");
    if (n_a + n_b == 0) return MSS_OK;
    MSS_REQUIRE((n_a == 0 || keys_a) && (n_b == 0 || keys_b) && workspace, "mss_sort_keys: null pointer");
    MSS_REQUIRE((((uintptr_t)keys_a | (uintptr_t)keys_b) & 3) == 0, "mss_sort_keys: keys must be 4-byte aligned");
    return sort_enqueue(nullptr, keys_a, n_a, keys_b, n_b, n_a + n_b, workspace, workspace_bytes, (cudaStream_t)stream,
                        nullptr);
}

static int keys_histogram(const uint32_t *keys, int64_t n, int bits, int every, int64_t *hist, void *stream, int refine = 0,
                          unsigned prefix = 0);

extern "C" int mss_keys_histogram(const uint32_t *keys, int64_t n, int bits, int64_t *hist, void *stream) {
    return keys_histogram(keys, n, bits, 1, hist, stream);
}

extern "C" int mss_keys_histogram_sampled(const uint32_t *keys, int64_t n, int bits, int every, int64_t *hist,
                                          void *stream) {
    MSS_REQUIRE(every >= 1, "mss_keys_histogram_sampled: every must be >= 1");
    return keys_histogram(keys, n, bits, every, hist, stream);
}

extern "C" int mss_keys_histogram_refine(const uint32_t *keys, int64_t n, unsigned prefix16, int every, int64_t *hist,
                                         void *stream) {
    MSS_REQUIRE(every >= 1 && prefix16 <= 0xffffu, "mss_keys_histogram_refine: every >= 1 and prefix16 < 65536 required");
    return keys_histogram(keys, n, 16, every, hist, stream, 1, prefix16);
}

static int keys_histogram(const uint32_t *keys, int64_t n, int bits, int every, int64_t *hist, void *stream, int refine,
                          unsigned prefix) {
    MSS_REQUIRE(bits >= 1 && bits <= 16 && hist, "mss_keys_histogram: bits must be 1..16");
    MSS_REQUIRE(n >= 0, "mss_keys_histogram: n < 0");
    cudaStream_t st = (cudaStream_t)stream;
    MSS_CHECK_CUDA(cudaMemsetAsync(hist, 0, sizeof(int64_t) << bits, st));
    if (n == 0) return MSS_OK;
    MSS_REQUIRE(keys, "mss_keys_histogram: null keys");
    MSS_CHECK_CUDA(opt_in_smem());
    // one CTA per SM (128 KB window each); small inputs use fewer CTAs (>= 16 K keys per CTA)
    int grid = (int)std::max<long long>(1, std::min<long long>((n / every + 16383) / 16384, (long long)sm_count()));
    keys_histogram_kernel<<<grid, KH_THREADS, KH_SMEM, st>>>(keys, n, 32 - bits, 1u << bits, every, (unsigned long long *)hist,
                                                            refine, prefix);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

// ---- key-range partition (multi-GPU exchange) --------------------------------------------------------------
static int status_stride(int parts) { return (parts + 15) / 16 * 16; }

extern "C" size_t mss_partition_workspace_bytes(int64_t n, int parts) {
    if (n < 0) n = 0;
    if (parts < 1) parts = 1;
    return 3 * RADIX * 8 + 256 + 256 + (sort_tiles(n) + 1) * (size_t)status_stride(parts) * 8 + 4096;
}

// parts <= 16 (one destination per GPU of a box): lane-private counters, no atomics in the loop (shared-memory atomics
// serialise 32-way when almost every key of a warp goes to the same destination); more parts: shared atomics.
constexpr int PC_SMALL = 16;
__global__ void __launch_bounds__(256)
partition_count_kernel(const uint32_t *__restrict__ keys, long long n, const uint32_t *__restrict__ splitters, int nspl,
                       int steps, unsigned long long *__restrict__ counts) {
    __shared__ unsigned s_c[PC_SMALL * 256];
    __shared__ uint32_t s_spl[RADIX];
    const bool small = nspl < PC_SMALL;
    for (int i = threadIdx.x; i < PC_SMALL * 256; i += 256) s_c[i] = 0;
    if ((int)threadIdx.x < nspl) s_spl[threadIdx.x] = __ldg(splitters + threadIdx.x);
    __syncthreads();
    const SplitterDigit dg{s_spl, nspl, steps};
    auto tally = [&](uint32_t k) {
        const unsigned d = dg(k);
        if (small) s_c[d * 256 + threadIdx.x]++;
        else atomicAdd(&s_c[d], 1u);
    };
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long head = min(n, (long long)(((16 - ((uintptr_t)keys & 15)) & 15) >> 2));
    const long long groups4 = (n - head) >> 2;
    const uint4 *k4 = reinterpret_cast<const uint4 *>(keys + head);
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; g + stride < groups4; g += 2 * stride) {                    // two 16-byte loads in flight per thread
        const uint4 a = __ldg(k4 + g), b = __ldg(k4 + g + stride);
        tally(a.x); tally(a.y); tally(a.z); tally(a.w);
        tally(b.x); tally(b.y); tally(b.z); tally(b.w);
    }
    if (g < groups4) {
        const uint4 a = __ldg(k4 + g);
        tally(a.x); tally(a.y); tally(a.z); tally(a.w);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = 0; i < n; i++) {
            if (i == head) i += groups4 << 2;
            if (i >= n) break;
            tally(__ldg(keys + i));
        }
    __syncthreads();
    if (small) {
        // warp w sums destinations w, w + 8
        const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int p = warp; p < PC_SMALL; p += 8) {
            unsigned long long acc = 0;
            for (int t = lane; t < 256; t += 32) acc += s_c[p * 256 + t];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0 && acc) atomicAdd(counts + p, acc);
        }
    } else if (s_c[threadIdx.x]) {
        atomicAdd(counts + threadIdx.x, (unsigned long long)s_c[threadIdx.x]);
    }
}

static int count_into(const uint32_t *keys, int64_t n, const uint32_t *splitters, int parts, unsigned long long *counts,
                      cudaStream_t st) {
    const int grid = (int)std::max<long long>(1, std::min<long long>((n + 2047) / 2048, (long long)sm_count() * 8));
    partition_count_kernel<<<grid, 256, 0, st>>>(keys, n, splitters, parts - 1, splitter_steps(parts), counts);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

extern "C" int mss_partition_count(const uint32_t *keys, int64_t n, const uint32_t *splitters, int parts,
                                   int64_t *out_counts_host, void *workspace, size_t workspace_bytes, void *stream) {
    MSS_REQUIRE(parts >= 1 && parts <= RADIX, "mss_partition_count: parts must be 1..256");
    MSS_REQUIRE(n >= 0 && out_counts_host, "mss_partition_count: bad arguments");
    for (int j = 0; j < parts; j++) out_counts_host[j] = 0;
    if (n == 0) return MSS_OK;
    MSS_REQUIRE(keys && workspace && workspace_bytes >= RADIX * 8 && (parts == 1 || splitters), "mss_partition_count: null pointer / workspace < 2048 bytes");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *counts = (unsigned long long *)workspace;
    MSS_CHECK_CUDA(cudaMemsetAsync(counts, 0, RADIX * 8, st));
    int rc = count_into(keys, n, splitters, parts, counts, st);
    if (rc) return rc;
    unsigned long long h[RADIX];
    MSS_CHECK_CUDA(cudaMemcpyAsync(h, counts, sizeof(h), cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    for (int j = 0; j < parts; j++) out_counts_host[j] = (int64_t)h[j];
    return MSS_OK;
}

// bucket d -> byte address table_host[d]; status / counter / device table carved from the workspace
static int scatter_to(const uint32_t *keys, int64_t n, const uint32_t *splitters, int parts,
                      const unsigned long long *table_host, void *workspace, size_t workspace_bytes, cudaStream_t st,
                      const char *who) {
    const size_t tiles = sort_tiles(n);
    const int sstride = status_stride(parts);
    Carver c(workspace, workspace_bytes);
    unsigned long long *table = c.take<unsigned long long>(RADIX);
    unsigned *counter = c.take<unsigned>(64);
    unsigned long long *status = c.take<unsigned long long>(tiles * sstride);
    if (!c.ok()) {
        set_error("%s: workspace too small (%zu < %zu)", who, workspace_bytes, mss_partition_workspace_bytes(n, parts));
        return MSS_ERR_WORKSPACE;
    }
    MSS_REQUIRE(tiles < (1ull << 31), "%s: n too large", who);
    MSS_CHECK_CUDA(cudaMemsetAsync(counter, 0, (size_t)((char *)(status + tiles * sstride) - (char *)counter), st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(table, table_host, (size_t)parts * 8, cudaMemcpyHostToDevice, st));
    static const bool no_bulk = [] { const char *e = getenv("MSS_PARTITION_NO_BULK"); return e && atoi(e); }();   // A/B switch
    if (parts <= PB_MAX_PARTS && !no_bulk)
        partition_scatter_bulk_kernel<<<(unsigned)tiles, SORT_THREADS, 0, st>>>(keys, n, splitters, parts - 1, splitter_steps(parts),
                                                                               table, status, sstride, counter);
    else
        partition_scatter_kernel<<<(unsigned)tiles, SORT_THREADS, 0, st>>>(keys, n, splitters, parts - 1, splitter_steps(parts),
                                                                          table, status, sstride, counter);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

extern "C" int mss_partition_scatter_keys(const uint32_t *keys, int64_t n, const uint32_t *splitters, int parts,
                                          const uint64_t *dst_keys_host, const int64_t *dst_offsets_host, void *workspace,
                                          size_t workspace_bytes, void *stream) {
    MSS_REQUIRE(parts >= 1 && parts <= RADIX, "mss_partition_scatter_keys: parts must be 1..256");
    MSS_REQUIRE(n >= 0 && dst_keys_host && dst_offsets_host, "mss_partition_scatter_keys: bad arguments");
    if (n == 0) return MSS_OK;
    MSS_REQUIRE(keys && workspace && (parts == 1 || splitters), "mss_partition_scatter_keys: null pointer");
    unsigned long long h[RADIX];
    for (int j = 0; j < parts; j++) {
        MSS_REQUIRE(dst_keys_host[j] && dst_offsets_host[j] >= 0 && (dst_keys_host[j] & 3) == 0,
                    "mss_partition_scatter_keys: bad destination %d", j);
        h[j] = (unsigned long long)dst_keys_host[j] + 4ull * (unsigned long long)dst_offsets_host[j];
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = scatter_to(keys, n, splitters, parts, h, workspace, workspace_bytes, st, "mss_partition_scatter_keys");
    if (rc) return rc;
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));     // the host staging array must outlive the async copy
    return MSS_OK;
}

extern "C" int mss_partition_keys(const uint32_t *keys, int64_t n, const uint32_t *splitters, int parts,
                                  uint32_t *keys_out, int64_t *out_counts_host, void *workspace, size_t workspace_bytes,
                                  void *stream) {
    MSS_REQUIRE(parts >= 1 && parts <= RADIX, "mss_partition_keys: parts must be 1..256");
    MSS_REQUIRE(n >= 0 && out_counts_host, "mss_partition_keys: bad arguments");
    for (int j = 0; j < parts; j++) out_counts_host[j] = 0;
    if (n == 0) return MSS_OK;
    MSS_REQUIRE(keys && keys_out && workspace && (parts == 1 || splitters), "mss_partition_keys: null pointer");
    MSS_REQUIRE(workspace_bytes >= mss_partition_workspace_bytes(n, parts), "mss_partition_keys: workspace too small");
    // the last 2 KB of the workspace hold the counts (scatter_to carves from the front)
    const size_t front = (workspace_bytes - 2048) & ~(size_t)255;
    int rc = mss_partition_count(keys, n, splitters, parts, out_counts_host, (char *)workspace + front, 2048, stream);
    if (rc) return rc;
    unsigned long long h[RADIX], run = 0;
    for (int j = 0; j < parts; j++) {
        h[j] = (unsigned long long)(uintptr_t)keys_out + 4ull * run;
        run += (unsigned long long)out_counts_host[j];
    }
    cudaStream_t st = (cudaStream_t)stream;
    rc = scatter_to(keys, n, splitters, parts, h, workspace, front, st, "mss_partition_keys");
    if (rc) return rc;
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    return MSS_OK;
}

// ---- both streams of an evaluator per call (one host synchronisation instead of two per exchange step) -------------
static bool eval_streams(const mss_eval_buffers *ev, int64_t n_neg, int64_t n_pos, const uint32_t **neg, const uint32_t **pos) {
    if (!ev || !ev->keys || n_neg < 0 || n_pos < 0 || n_neg + n_pos > ev->capacity) return false;
    *neg = ev->keys;
    *pos = ev->keys + (ev->capacity - n_pos);
    return true;
}

extern "C" size_t mss_eval_partition_workspace_bytes(int64_t n_neg, int64_t n_pos, int parts) {
    return 4096 + 2 * RADIX * 8 + align_up(mss_partition_workspace_bytes(n_neg, parts), 256) +
           align_up(mss_partition_workspace_bytes(n_pos, parts), 256);
}

extern "C" int mss_eval_partition_count(const mss_eval_buffers *ev, int64_t n_neg, int64_t n_pos, const uint32_t *splitters_host,
                                        int parts, int64_t *out_counts_host, void *workspace, size_t workspace_bytes,
                                        void *stream) {
    MSS_REQUIRE(parts >= 1 && parts <= RADIX && out_counts_host && workspace, "mss_eval_partition_count: bad arguments");
    MSS_REQUIRE(parts == 1 || splitters_host, "mss_eval_partition_count: null splitters");
    const uint32_t *neg, *pos;
    MSS_REQUIRE(eval_streams(ev, n_neg, n_pos, &neg, &pos), "mss_eval_partition_count: bad evaluator / stream sizes");
    MSS_REQUIRE(workspace_bytes >= 4096 + 2 * RADIX * 8, "mss_eval_partition_count: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *spl = (uint32_t *)workspace;                                      // [256]
    unsigned long long *counts = (unsigned long long *)((char *)workspace + 4096);   // [2][256]
    if (parts > 1) MSS_CHECK_CUDA(cudaMemcpyAsync(spl, splitters_host, (size_t)(parts - 1) * 4, cudaMemcpyHostToDevice, st));
    MSS_CHECK_CUDA(cudaMemsetAsync(counts, 0, 2 * RADIX * 8, st));
    int rc = MSS_OK;
    if (n_neg) rc = count_into(neg, n_neg, spl, parts, counts, st);
    if (rc == MSS_OK && n_pos) rc = count_into(pos, n_pos, spl, parts, counts + RADIX, st);
    if (rc) return rc;
    unsigned long long h[2 * RADIX];
    MSS_CHECK_CUDA(cudaMemcpyAsync(h, counts, sizeof(h), cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    for (int j = 0; j < parts; j++) {
        out_counts_host[j] = (int64_t)h[j];
        out_counts_host[parts + j] = (int64_t)h[RADIX + j];
    }
    return MSS_OK;
}

extern "C" int mss_eval_partition_scatter(const mss_eval_buffers *ev, int64_t n_neg, int64_t n_pos,
                                          const uint32_t *splitters_host, int parts, const uint64_t *dst_keys_host,
                                          const int64_t *dst_neg_offsets_host, const int64_t *dst_pos_offsets_host,
                                          void *workspace, size_t workspace_bytes, void *stream) {
    MSS_REQUIRE(parts >= 1 && parts <= RADIX && dst_keys_host && dst_neg_offsets_host && dst_pos_offsets_host && workspace,
                "mss_eval_partition_scatter: bad arguments");
    MSS_REQUIRE(parts == 1 || splitters_host, "mss_eval_partition_scatter: null splitters");
    const uint32_t *neg, *pos;
    MSS_REQUIRE(eval_streams(ev, n_neg, n_pos, &neg, &pos), "mss_eval_partition_scatter: bad evaluator / stream sizes");
    if (workspace_bytes < mss_eval_partition_workspace_bytes(n_neg, n_pos, parts)) {
        set_error("mss_eval_partition_scatter: workspace too small (%zu < %zu)", workspace_bytes,
                  mss_eval_partition_workspace_bytes(n_neg, n_pos, parts));
        return MSS_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *spl = (uint32_t *)workspace;
    if (parts > 1) MSS_CHECK_CUDA(cudaMemcpyAsync(spl, splitters_host, (size_t)(parts - 1) * 4, cudaMemcpyHostToDevice, st));
    unsigned long long hn[RADIX], hp[RADIX];
    for (int j = 0; j < parts; j++) {
        MSS_REQUIRE(dst_keys_host[j] && (dst_keys_host[j] & 3) == 0 && dst_neg_offsets_host[j] >= 0 && dst_pos_offsets_host[j] >= 0,
                    "mss_eval_partition_scatter: bad destination %d", j);
        hn[j] = (unsigned long long)dst_keys_host[j] + 4ull * (unsigned long long)dst_neg_offsets_host[j];
        hp[j] = (unsigned long long)dst_keys_host[j] + 4ull * (unsigned long long)dst_pos_offsets_host[j];
    }
    char *w0 = (char *)workspace + 4096 + 2 * RADIX * 8;
    const size_t b0 = align_up(mss_partition_workspace_bytes(n_neg, parts), 256);
    int rc = MSS_OK;
    if (n_neg) rc = scatter_to(neg, n_neg, spl, parts, hn, w0, b0, st, "mss_eval_partition_scatter");
    if (rc == MSS_OK && n_pos)
        rc = scatter_to(pos, n_pos, spl, parts, hp, w0 + b0, align_up(mss_partition_workspace_bytes(n_pos, parts), 256), st,
                        "mss_eval_partition_scatter");
    if (rc) return rc;
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));     // the host staging arrays must outlive the async copies
    return MSS_OK;
}

// Exchange by remote append: every key of this rank's evaluator is appended to the evaluator of the rank that owns its key
// range -- dst_keys_host[d] / dst_state_host[d] are the device addresses of rank d's key buffer and state (peer-mapped
// memory for d != this rank), each of dst_capacity keys.  One launch, no sizes needed in advance.  The caller zeroes the
// destination states, orders the call between two barriers, and afterwards every rank reads its own state
// (mss_eval_state_host reports a capacity overflow).  parts <= 32.  Synchronises the stream.
extern "C" int mss_eval_exchange_append(const mss_eval_buffers *ev, int64_t n_neg, int64_t n_pos, const uint32_t *splitters_host,
                                        int parts, const uint64_t *dst_keys_host, const uint64_t *dst_state_host,
                                        int64_t dst_capacity, void *workspace, size_t workspace_bytes, void *stream) {
    MSS_REQUIRE(parts >= 1 && parts <= PB_MAX_PARTS, "mss_eval_exchange_append: parts must be 1..%d", PB_MAX_PARTS);
    MSS_REQUIRE(dst_keys_host && dst_state_host && dst_capacity > 0 && workspace && workspace_bytes >= 4096,
                "mss_eval_exchange_append: bad arguments");
    MSS_REQUIRE(parts == 1 || splitters_host, "mss_eval_exchange_append: null splitters");
    const uint32_t *neg, *pos;
    MSS_REQUIRE(eval_streams(ev, n_neg, n_pos, &neg, &pos), "mss_eval_exchange_append: bad evaluator / stream sizes");
    if (n_neg + n_pos == 0) return MSS_OK;
    ExchangeDst dst;
    for (int j = 0; j < PB_MAX_PARTS; j++) { dst.keys[j] = 0; dst.state[j] = 0; }
    for (int j = 0; j < parts; j++) {
        MSS_REQUIRE(dst_keys_host[j] && dst_state_host[j] && (dst_keys_host[j] & 15) == 0 && (dst_state_host[j] & 7) == 0,
                    "mss_eval_exchange_append: bad destination %d", j);
        dst.keys[j] = dst_keys_host[j];
        dst.state[j] = dst_state_host[j];
    }
    dst.capacity = dst_capacity;
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *spl = (uint32_t *)workspace;
    if (parts > 1) MSS_CHECK_CUDA(cudaMemcpyAsync(spl, splitters_host, (size_t)(parts - 1) * 4, cudaMemcpyHostToDevice, st));
    const size_t tn = sort_tiles(n_neg), tp = sort_tiles(n_pos);
    MSS_REQUIRE(tn + tp < (1ull << 31), "mss_eval_exchange_append: too many keys");
    exchange_append_kernel<<<(unsigned)(tn + tp), SORT_THREADS, 0, st>>>(neg, n_neg, pos, n_pos, (unsigned)tn, spl, parts - 1,
                                                                        splitter_steps(parts), dst);
    MSS_CHECK_LAUNCH();
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    return MSS_OK;
}

// Streaming form: enqueue (no host synchronisation) the exchange of a STAGING evaluator -- its sizes are read from its
// device state -- followed by `accum_state += staging state; staging state = 0`.  splitters_dev is a DEVICE array.
// With a workspace (mss_eval_exchange_stream_workspace_bytes) and parts <= 16 the counted form runs: per-destination
// counts of the batch, ONE remote reservation per destination and stream, look-back / bulk-store scatter; without one,
// every tile reserves its own runs (exchange_append_dev_kernel).
extern "C" size_t mss_eval_exchange_stream_workspace_bytes(int64_t staging_capacity, int parts) {
    if (staging_capacity < 0) staging_capacity = 0;
    const size_t tiles = sort_tiles(staging_capacity) + 2;
    return 2 * RADIX * 8 + 2 * PB_MAX_PARTS * 8 + 256 + 2 * tiles * (size_t)status_stride(parts) * 8 + 2048;
}

extern "C" int mss_eval_exchange_stream(const mss_eval_buffers *staging, const uint32_t *splitters_dev, int parts,
                                        const uint64_t *dst_keys_host, const uint64_t *dst_state_host, int64_t dst_capacity,
                                        void *accum_state, void *workspace, size_t workspace_bytes, void *stream) {
    MSS_REQUIRE(parts >= 1 && parts <= PB_MAX_PARTS, "mss_eval_exchange_stream: parts must be 1..%d", PB_MAX_PARTS);
    MSS_REQUIRE(staging && staging->keys && staging->state && staging->capacity > 0 && accum_state && dst_keys_host &&
                    dst_state_host && dst_capacity > 0 && (parts == 1 || splitters_dev),
                "mss_eval_exchange_stream: bad arguments");
    ExchangeDst dst;
    for (int j = 0; j < PB_MAX_PARTS; j++) { dst.keys[j] = 0; dst.state[j] = 0; }
    for (int j = 0; j < parts; j++) {
        MSS_REQUIRE(dst_keys_host[j] && dst_state_host[j] && (dst_keys_host[j] & 15) == 0 && (dst_state_host[j] & 7) == 0,
                    "mss_eval_exchange_stream: bad destination %d", j);
        dst.keys[j] = dst_keys_host[j];
        dst.state[j] = dst_state_host[j];
    }
    dst.capacity = dst_capacity;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t tiles = sort_tiles(staging->capacity) + 2;            // two streams: up to one partial tile each
    MSS_REQUIRE(tiles < (1ull << 31), "mss_eval_exchange_stream: staging buffer too large");
    const int steps = splitter_steps(parts);
    if (workspace && parts <= 16) {
        if (workspace_bytes < mss_eval_exchange_stream_workspace_bytes(staging->capacity, parts)) {
            set_error("mss_eval_exchange_stream: workspace too small (%zu < %zu)", workspace_bytes,
                      mss_eval_exchange_stream_workspace_bytes(staging->capacity, parts));
            return MSS_ERR_WORKSPACE;
        }
        const int sstride = status_stride(parts);
        Carver c(workspace, workspace_bytes);
        unsigned long long *counts = c.take<unsigned long long>(2 * RADIX);
        unsigned long long *table = c.take<unsigned long long>(2 * PB_MAX_PARTS);
        unsigned *counters = c.take<unsigned>(64);
        unsigned long long *status = c.take<unsigned long long>(2 * tiles * sstride);
        MSS_REQUIRE(c.ok(), "mss_eval_exchange_stream: internal workspace layout");
        MSS_CHECK_CUDA(cudaMemsetAsync(workspace, 0, (size_t)((char *)(status + 2 * tiles * sstride) - (char *)workspace), st));
        const StagingView sv{staging->keys, staging->capacity, (const EvalState *)staging->state};
        const int cgrid = (int)std::max<long long>(1, std::min<long long>((staging->capacity + 4095) / 4096, (long long)sm_count() * 4));
        partition_count_dev_kernel<<<dim3(cgrid, 2), 256, 0, st>>>(sv, splitters_dev, parts - 1, steps, counts);
        MSS_CHECK_LAUNCH();
        exchange_reserve_kernel<<<1, 2 * PB_MAX_PARTS, 0, st>>>(counts, dst, parts, table);
        MSS_CHECK_LAUNCH();
        partition_scatter_bulk_dev_kernel<<<dim3((unsigned)tiles, 2), SORT_THREADS, 0, st>>>(sv, splitters_dev, parts - 1, steps, table,
                                                                                           status, tiles * sstride, sstride, counters);
        MSS_CHECK_LAUNCH();
    } else {
        // MSS_EXCHANGE_CTAS_PER_SM (default 1): resident CTAs of the streamed exchange per SM; 0 = one CTA per tile
        static const int per_sm = [] { const char *e = getenv("MSS_EXCHANGE_CTAS_PER_SM"); return e ? atoi(e) : 1; }();
        const unsigned grid = per_sm > 0 ? (unsigned)std::min<size_t>(tiles, (size_t)sm_count() * per_sm) : (unsigned)tiles;
        exchange_append_dev_kernel<<<grid, SORT_THREADS, 0, st>>>(staging->keys, staging->capacity,
                                                                            (const EvalState *)staging->state, splitters_dev,
                                                                            parts - 1, steps, dst);
        MSS_CHECK_LAUNCH();
    }
    eval_state_fold_kernel<<<1, 1, 0, st>>>((EvalState *)accum_state, (EvalState *)staging->state);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

extern "C" int mss_eval_exchange_stage(const mss_eval_buffers *staging, const uint32_t *splitters_dev, int parts,
                                       const uint64_t *outbox_keys_host, const uint64_t *outbox_state_host,
                                       int64_t outbox_capacity, const uint64_t *recv_state_host, int64_t recv_capacity,
                                       void *accum_state, uint64_t *plan, void *stream) {
    MSS_REQUIRE(parts >= 1 && parts <= PB_MAX_PARTS, "mss_eval_exchange_stage: parts must be 1..%d", PB_MAX_PARTS);
    MSS_REQUIRE(staging && staging->keys && staging->state && staging->capacity > 0 && accum_state && outbox_keys_host &&
                    outbox_state_host && recv_state_host && plan && outbox_capacity > 0 && recv_capacity > 0 &&
                    (parts == 1 || splitters_dev),
                "mss_eval_exchange_stage: bad arguments");
    ExchangeDst box, recv;
    for (int j = 0; j < PB_MAX_PARTS; j++) { box.keys[j] = box.state[j] = recv.keys[j] = recv.state[j] = 0; }
    for (int j = 0; j < parts; j++) {
        MSS_REQUIRE(outbox_keys_host[j] && outbox_state_host[j] && recv_state_host[j] && (outbox_keys_host[j] & 15) == 0 &&
                        (outbox_state_host[j] & 7) == 0 && (recv_state_host[j] & 7) == 0,
                    "mss_eval_exchange_stage: bad buffer %d", j);
        box.keys[j] = outbox_keys_host[j];
        box.state[j] = outbox_state_host[j];
        recv.state[j] = recv_state_host[j];
    }
    box.capacity = outbox_capacity;
    recv.capacity = recv_capacity;
    cudaStream_t st = (cudaStream_t)stream;
    bool packed = true;                                                 // states back to back: one memset instead of `parts`
    for (int j = 1; j < parts; j++) packed &= outbox_state_host[j] == outbox_state_host[j - 1] + MSS_EVAL_STATE_BYTES;
    if (packed) {
        MSS_CHECK_CUDA(cudaMemsetAsync((void *)outbox_state_host[0], 0, (size_t)parts * MSS_EVAL_STATE_BYTES, st));
    } else {
        for (int j = 0; j < parts; j++) MSS_CHECK_CUDA(cudaMemsetAsync((void *)outbox_state_host[j], 0, MSS_EVAL_STATE_BYTES, st));
    }
    const size_t tiles = sort_tiles(staging->capacity) + 2;            // two streams: up to one partial tile each
    MSS_REQUIRE(tiles < (1ull << 31), "mss_eval_exchange_stage: staging buffer too large");
    exchange_append_dev_kernel<<<(unsigned)tiles, SORT_THREADS, 0, st>>>(staging->keys, staging->capacity,
                                                                        (const EvalState *)staging->state, splitters_dev,
                                                                        parts - 1, splitter_steps(parts), box);
    MSS_CHECK_LAUNCH();
    exchange_plan_kernel<<<1, 2 * PB_MAX_PARTS, 0, st>>>(box, recv, parts, (unsigned long long *)plan);
    MSS_CHECK_LAUNCH();
    eval_state_fold_kernel<<<1, 1, 0, st>>>((EvalState *)accum_state, (EvalState *)staging->state);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

extern "C" int mss_memcpy_async(void *dst, const void *src, size_t bytes, void *stream) {
    if (bytes == 0) return MSS_OK;
    MSS_REQUIRE(dst && src, "mss_memcpy_async: null pointer");
    MSS_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return MSS_OK;
}
