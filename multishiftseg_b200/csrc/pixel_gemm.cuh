// Persistent "pixels x channels" GEMM on tcgen05 shared by the two head fusions of SURVEY 8f-1:
//   * mask_gemm.cu -- Mask2Former mask logits, einsum("bqc,bchw->bqhw") (N = 112 queries, per-image B table)
//   * head_gemm.cu -- DeepLabv3+ final / ood_head 1x1 convolutions + energy (N = 48, one B table)
//     D[px, n] = sum_k F[k, px] W[n, k],   F NCHW fp32 (pixel-contiguous), M = 128 pixels per tile, K = 32 * STAGES
// 3xTF32 with fp32 accumulation in TMEM (F = F_hi + F_lo, W = W_hi + W_lo; hi = top 19 bits, lo = exact remainder):
//     D = F_lo*W_hi + F_hi*W_lo + F_hi*W_hi
// The feature map is "M-major"; instead of staging it through a swizzled shared-memory layout the producer warps read
// it with plain coalesced loads (lane <-> pixel <-> TMEM lane), split it in registers and write it straight into
// TENSOR MEMORY as the A operand (tcgen05.st); the pre-split weights are the B operand in shared memory (K-major core
// matrices, no swizzle; up to 224 KB, so there is no shared memory to stage features in -- they wait in REGISTERS):
//   * 8 producer warps + 1 MMA-issuer warp (288 threads = 12 allocated warps -> 168 registers each), one CTA per SM;
//     warp = (lane quarter, half): stage = 32 channels, a thread's half takes 16 of them;
//   * a thread keeps a ring of STAGES x 16 loaded values = one whole tile (128 registers for K = 256): right after
//     stage s of tile t has been split and stored to TMEM, the loads of stage s of tile t+1 are issued into the same
//     registers, so every load has a full tile time to arrive and ~120 KB per SM are in flight;
//   * 4 A slots in TMEM (64 columns each: 32 hi + 32 lo), two accumulators (at columns 0 and 128) -> 512 columns;
//   * work item = (image, slice): a CTA walks tiles slice, slice + S, ... of one image (S = SMs / B slices per image
//     when B <= SMs: one item per CTA); a per-image B table arrives as ONE bulk copy (cp.async.bulk -> mbarrier)
//     issued by the MMA warp once the tensor work of the previous item is complete; the producers never touch it
//     and prefetch straight across item boundaries;
//   * epilogue of tile t: Epi::CHUNKS chunks, spread over the stages of tile t+1 (policy class, see the two users).
// Addresses: ONE running pointer per tile, advanced by a plane per load and pinned with an empty asm (ptxas otherwise
// precomputes dozens of 64-bit addresses and spills freshly loaded values); warp index through a shuffle so that
// ptxas keeps the role branches and the load descriptors uniform.
//
// History (round 1, B200, mask GEMM on the cfg-3 batch 8 x [100 x 256] x [256 x 256 x 512]):
//   v1  two 8-warp pipelines per CTA, 2-deep register buffer, 96 registers, `live ? ld(src + i*hw) : 0` addressing:
//       0.370 ms; ncu: 21 warp instructions per element (8 of them predicated 64-bit address arithmetic per load), 41 %
//       of the stall samples on the first use of a loaded value, DRAM 47 %, tensor pipe 50 %
//   v1 + running pointers + uniform warp index: 0.357 ms (the same change took head_gemm 1.19 -> 0.97 ms)
//   v1 + L2 prefetch of the next tile (prefetch.global.L2 per 128-byte line): 4 % SLOWER, dropped
//   v2  this structure: 0.327 ms = 4.57 TB/s (70 % of the HBM peak); ncu: 72 M instead of 176 M instructions, DRAM 56 %,
//       tensor pipe 61 % (its floor: 3 MMAs x N = 112 -> 0.19 ms; the HBM floor is 0.23 ms); ring depth 6 vs 8: same.
#pragma once
#include <type_traits>

#include "tc5_common.cuh"

namespace mss {

constexpr int PG_STAGE_K = 32;
constexpr int PG_PRODUCERS = 256;
constexpr int PG_THREADS = PG_PRODUCERS + 32;
constexpr int PG_SLOTS = 4;
constexpr int PG_TMEM_COLS = 512;
constexpr int PG_COL_A = 256;                      // accumulators at columns 0 and 128, A slot k at 256 + 64 k
constexpr int PG_MAX_K = 256;
template <int N>
struct PgIdesc {
    static constexpr uint32_t value = tc5_idesc_tf32(128, N);     // evaluated on the host side of the compiler
};

// the CTA's flat tile sequence: items blockIdx.x, + gridDim.x, ...; inside item (image b, slice): tiles slice + j * slices
struct TileCursor {
    long long item, n_items, b;
    int slices, tiles_per_image, slice, j, n_j;
    __device__ __forceinline__ void set_item() {
        b = item / slices;
        slice = (int)(item - b * slices);
        n_j = (tiles_per_image - slice + slices - 1) / slices;      // >= 1: the host keeps slices <= tiles_per_image
        j = 0;
    }
    __device__ __forceinline__ TileCursor(long long first, long long n_items_, int slices_, int tpi)
        : item(first), n_items(n_items_), b(0), slices(slices_), tiles_per_image(tpi), slice(0), j(0), n_j(0) {
        if (item < n_items) set_item();
    }
    __device__ __forceinline__ bool valid() const { return item < n_items; }
    __device__ __forceinline__ long long first_pixel() const { return ((long long)slice + (long long)j * slices) * 128; }
    __device__ __forceinline__ void next() {
        if (++j >= n_j) {
            item += gridDim.x;
            if (item < n_items) set_item();
        }
    }
};

// Epi (epilogue policy): `static constexpr int CHUNKS`; `struct Tile` (per-thread output state of one tile);
//   begin(Tile &, image b, first pixel p of this thread's row or -1 when it lies past the end, half)
//   chunk(Tile &, c, taddr of column 0 of the tile's accumulator in this thread's lane, half, plane bytes)
// `table_stride`: floats between consecutive images' B tables (0: one table for all images).
template <int STAGES, int N, class Epi>
__global__ void __launch_bounds__(PG_THREADS, 1)
pixel_gemm_kernel(const float *__restrict__ feat, int hw, long long n_items, int slices, int tiles_per_image,
                  const float *__restrict__ table, long long table_stride, const Epi epi) {
    constexpr int K = STAGES * PG_STAGE_K;
    constexpr int CHUNKS = Epi::CHUNKS;
    constexpr int CPS = (CHUNKS + STAGES - 1) / STAGES;                          // epilogue chunks per stage
    constexpr uint32_t IDESC = PgIdesc<N>::value;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_b = reinterpret_cast<float *>(smem_raw);                            // [hi | lo][K/4][N][4]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_b + 2 * K * N);          // full[4] empty[4] dfull[2] dempty[2] table
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 16);
    uint64_t *bar_full = s_bar, *bar_empty = s_bar + 4, *bar_dfull = s_bar + 8, *bar_dempty = s_bar + 10, *bar_table = s_bar + 12;

    const int tid = threadIdx.x, lane = tid & 31;
    // warp index through a shuffle: ptxas then knows it is warp-uniform, keeps the role branches uniform (BRA.U) and
    // the load descriptors / loop state in uniform registers instead of re-materialising them (R2UR) at every load
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    constexpr unsigned table_bytes = 2u * K * N * 4u;

    if (tid == PG_PRODUCERS) {
        for (int k = 0; k < PG_SLOTS; k++) { mbar_init(&bar_full[k], PG_PRODUCERS); mbar_init(&bar_empty[k], 1); }
        for (int k = 0; k < 2; k++) { mbar_init(&bar_dfull[k], 1); mbar_init(&bar_dempty[k], PG_PRODUCERS); }
        mbar_init(bar_table, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "n"(PG_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *s_tmem;

    if (warp < 8) {
        // ===== producers + epilogue =====
        const int quarter = warp & 3, half = warp >> 2;
        const int m = quarter * 32 + lane;                                       // pixel inside the tile == TMEM lane
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
        const size_t plane = (size_t)(unsigned)hw * sizeof(float);

        // Register ring of RING stage buffers: the loads of flat stage k + RING are issued right after stage k has been
        // consumed, into the same registers.  For K = 256 (8 stages per tile) RING = 6: ~88 loads per thread in flight,
        // three quarters of a tile ahead; the ring position of a tile's first stage cycles through 0, 2, 4, so the tile
        // loop is unrolled three times (static register indices).  (RING = 8 needs 183 registers; 288 threads are
        // allocated as 12 warps, i.e. 168 registers per thread at most -- it spilled freshly loaded values.)
        constexpr int RING = STAGES;                                            // (the loop below also supports RING < STAGES)
        constexpr int LAG = STAGES - RING;                                      // stages of the SAME tile still to prefetch
        float buf[RING][16];
        const char *qn = nullptr;                                                // running load pointer (tile being prefetched)
        typename Epi::Tile o_cur, o_prev, o_next;                                // per-tile output state of the epilogue policy
        unsigned u = 0, t = 0;                                                   // stage uses / tiles so far (only parities matter)
        bool have_next = false, have_prev = false;

        TileCursor pf(blockIdx.x, n_items, slices, tiles_per_image);
        // thread's load pointer and output pointer for the cursor's tile
        auto setup = [&](const TileCursor &c, typename Epi::Tile &o) {
            const long long p = c.first_pixel() + m;
            const long long pc = p < hw ? p : (long long)hw - 1;                 // rows past the end read the last pixel
            qn = reinterpret_cast<const char *>(feat + (c.b * K + half * 16) * (long long)hw + pc);
            epi.begin(o, c.b, p < hw ? p : -1, half);
        };
        auto load_stage = [&](float (&dst)[16]) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                dst[i] = ldg_stream_f1(reinterpret_cast<const float *>(qn));
                qn += plane;
                asm volatile("" : "+l"(qn));     // keep ONE running pointer (ptxas otherwise precomputes and spills dozens)
            }
            qn += 16 * plane;                                                    // the other half's channels
            asm volatile("" : "+l"(qn));
        };
        // chunk c of the epilogue of tile t-1 (accumulator (t-1) & 1)
        auto epi_chunk = [&](int c) {
            const unsigned tp = t - 1, db = tp & 1;
            if (c == 0) {
                mbar_wait(&bar_dfull[db], (tp >> 1) & 1);
                tc5_fence_after();
            }
            epi.chunk(o_prev, c, lane_base + db * 128, half, plane);
            if (c == CHUNKS - 1) {
                tc5_fence_before();
                mbar_arrive(&bar_dempty[db]);                                    // the accumulator may be overwritten
            }
        };
        // one tile whose first stage sits at ring position OFF
        auto tile_body = [&](auto off_c) {
            constexpr int OFF = decltype(off_c)::value;
#pragma unroll
            for (int s = 0; s < STAGES; s++, u++) {
                float (&cur)[16] = buf[(OFF + s) % RING];
                const unsigned slot = u & (PG_SLOTS - 1);
                if (u >= PG_SLOTS) mbar_wait(&bar_empty[slot], ((u >> 2) + 1) & 1);   // MMAs of use u-4 are done
                tc5_fence_after();
                const uint32_t a = lane_base + PG_COL_A + slot * 64 + half * 16;
#pragma unroll
                for (int c = 0; c < 2; c++) {                                    // 8 channels at a time: 16 temporaries
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        hi[i] = __float_as_uint(cur[8 * c + i]) & 0xFFFFE000u;
                        lo[i] = __float_as_uint(cur[8 * c + i] - __uint_as_float(hi[i]));
                    }
                    tc5_st8(a + 8 * c, hi);
                    tc5_st8(a + 32 + 8 * c, lo);
                }
                tc5_wait_st();
                tc5_fence_before();
                mbar_arrive(&bar_full[slot]);
                // flat stage + RING: the rest of this tile first, then the next tile
                if (s < LAG) load_stage(cur);
                else {
                    if (s == LAG) {
                        have_next = pf.valid();
                        if (have_next) { setup(pf, o_next); pf.next(); }
                    }
                    if (have_next) load_stage(cur);
                }
                if (have_prev) {
#pragma unroll
                    for (int c = s * CPS; c < (s + 1) * CPS && c < CHUNKS; c++) epi_chunk(c);
                }
            }
            t++;
            o_prev = o_cur;
            have_prev = true;
            o_cur = o_next;
        };

        bool have = pf.valid();
        if (have) {
            setup(pf, o_cur);
            pf.next();
#pragma unroll
            for (int s = 0; s < RING; s++) load_stage(buf[s]);
        }
        while (have) {
            tile_body(std::integral_constant<int, 0>{});
            have = have_next;
            if (STAGES % RING != 0) {                                            // K = 256: ring positions 0, 2, 4
                if (!have) break;
                tile_body(std::integral_constant<int, (STAGES) % RING>{});
                have = have_next;
                if (!have) break;
                tile_body(std::integral_constant<int, (2 * STAGES) % RING>{});
                have = have_next;
            }
        }
        if (have_prev) {
#pragma unroll
            for (int c = 0; c < CHUNKS; c++) epi_chunk(c);
        }
    } else {
        // ===== MMA issuer: the whole warp waits (stays converged), one elected lane issues =====
        const uint32_t bhi = smem_u32(s_b), blo = bhi + (uint32_t)K * N * 4;
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);               // warp-uniform (see m2f_tc5q.cuh)
        unsigned u = 0, t = 0, n_loaded = 0;
        long long cur_item = -1;
        for (TileCursor c(blockIdx.x, n_items, slices, tiles_per_image); c.valid(); c.next(), t++) {
            if (c.item != cur_item && (table_stride != 0 || cur_item < 0)) {
                // the image's table: every MMA that read the previous one has completed (commits are in order)
                if (t > 0) mbar_wait_backoff(&bar_dfull[(t - 1) & 1], ((t - 1) >> 1) & 1, 32);
                if (elect_one_sync()) {
                    mbar_expect_tx(bar_table, table_bytes);
                    bulk_load_1d(s_b, table + c.b * table_stride, table_bytes, bar_table);
                }
                __syncwarp();
                mbar_wait_backoff(bar_table, n_loaded & 1, 32);
                n_loaded++;
                cur_item = c.item;
            }
            const unsigned db = t & 1;
            if (t >= 2) mbar_wait_backoff(&bar_dempty[db], ((t >> 1) + 1) & 1, 32);   // epilogue of tile t-2 has read this accumulator
            const uint32_t d = tmem_u + db * 128;
#pragma unroll 1
            for (int s = 0; s < STAGES; s++, u++) {
                const unsigned slot = u & (PG_SLOTS - 1);
                mbar_wait_backoff(&bar_full[slot], (u >> 2) & 1, 32);
                tc5_fence_after();
                if (elect_one_sync()) {
#pragma unroll
                    for (int kk = 0; kk < PG_STAGE_K / 8; kk++) {
                        const int ks = s * (PG_STAGE_K / 8) + kk;                // k-step of 8 channels = 2 core-matrix chunks
                        const uint64_t dh = tc5_smem_desc(bhi + ks * 2 * (N * 16), N * 16, 128);
                        const uint64_t dl = tc5_smem_desc(blo + ks * 2 * (N * 16), N * 16, 128);
                        const uint32_t ahi = tmem_u + PG_COL_A + slot * 64 + kk * 8, alo = ahi + 32;
                        tc5_mma_ts(d, alo, dh, IDESC, (s | kk) > 0);
                        tc5_mma_ts(d, ahi, dl, IDESC, 1);
                        tc5_mma_ts(d, ahi, dh, IDESC, 1);
                    }
                    tc5_commit(&bar_empty[slot]);
                    if (s == STAGES - 1) tc5_commit(&bar_dfull[db]);
                }
                __syncwarp();
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc5_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(PG_TMEM_COLS) : "memory");
    }
}

constexpr size_t pixel_gemm_smem(int K, int N) { return (size_t)2 * K * N * 4 + 16 * 8 + 16; }

// slices per image / items / grid for B images of `tiles_per_image` tiles on `sms` SMs
struct PixelGemmPlan {
    int slices, grid;
    long long n_items;
};
inline PixelGemmPlan pixel_gemm_plan(long long B, int tiles_per_image, int sms, bool shared_table = false) {
    PixelGemmPlan p;
    // B <= SMs: S = SMs / B slices per image, one work item per CTA; otherwise whole images, round-robin
    p.slices = (B <= sms) ? (tiles_per_image < sms / (int)B ? tiles_per_image : sms / (int)B) : 1;
    if (p.slices < 1) p.slices = 1;
    if (shared_table && B <= sms) {
        // One table for every image (the DeepLab head): items need not be image-aligned per CTA, so pick the slice count
        // that makes B * S a multiple of the SM count -- every SM gets the same number of items (B = 8: S = 37, 2 items
        // per CTA on all 148 SMs instead of 1 item on 144 of them).
        long long a = B, b = sms;
        while (b) { const long long t = a % b; a = b; b = t; }          // a = gcd(B, sms)
        const int unit = (int)(sms / a);
        const int want = (int)((sms + B - 1) / B);
        const int s = (want + unit - 1) / unit * unit;
        if (s <= tiles_per_image) p.slices = s;
    }
    p.n_items = B * p.slices;
    p.grid = (int)(p.n_items < sms ? p.n_items : sms);
    return p;
}

}  // namespace mss
