// Mask2Former fused post-head inference, tcgen05 variant (exact x4 upsample path) -- the default fast path.
//
// Same contract as m2f_fused_x4_kernel / m2f_mma_x4_kernel.  ncu on the two earlier variants (round 1):
// FFMA contraction -> fma pipe 54 %, issue 66 %; mma.sync 3xTF32 -> legacy tensor pipe 43 % busy at
// 277 us/image, i.e. a 120 us/image floor on that path alone, with the HMMAs sharing issue slots with the
// interpolation + sigmoid work.  Here the 100 x 19 contraction
//     semseg[px, c] = sum_q S[px, q] P[q, c]
// runs on the 5th-generation tensor cores as a 3xTF32 split GEMM (tcgen05.mma kind::tf32, fp32 accumulate
// in TMEM), issued by ONE thread, so the CUDA cores only interpolate, evaluate the sigmoid and split:
//     S = S_hi + S_lo,  P = P_hi + P_lo  (hi = top 19 bits, lo = exact remainder)
//     D  = S_lo*P_hi;  D += S_hi*P_lo;  D += S_hi*P_hi          (S_lo*P_lo ~ 2^-22, dropped)
//
// Geometry: CTA = 64 x 16 output pixels of one image = 8 M-tiles of 64 x 2 px (M = 128), N = 32 (19 classes
// padded), K = 104 queries (13 k-steps of 8).  One TMA box (24 x 6 x 104 fp32, 59.9 KB) brings the low-res
// patch of ALL queries for the CTA; the zero-filled out-of-image halo is overwritten with the edge value
// (torch clamps source indices), after which every tap is border-free.
//   * warps 0-7 (producers): thread <-> one pixel row of the M-tile (= one TMEM lane) and one half of the
//     queries of a 16-query stage: 4 LDS + 4 FMA-pipe ops + ex2 + add + rcp + split per (pixel, query),
//     written straight into TENSOR MEMORY as the A operand (tcgen05.st 32x32b.x8: hi and lo), never through
//     shared memory;
//   * warp 8, one lane (MMA issuer): per stage waits for the 256 producer arrivals, issues 3 x 2
//     tcgen05.mma (A from TMEM, B = class-probability table in shared memory, K-major, no swizzle),
//     tcgen05.commit releases the stage / publishes the accumulator;
//   * warps 0-3 (epilogue, one tile behind): tcgen05.ld of the 128 x 32 fp32 accumulator, 1 - max_c and/or
//     the 19 class planes, coalesced 128-byte row stores.
// TMEM: A_hi cols 0..103, A_lo cols 104..207, D cols 208..239 -> 256 columns per CTA, two CTAs per SM.
#pragma once

#include "tc5_common.cuh"

namespace mss {

constexpr int T5_TILE_W = 64, T5_BLOCK_H = 16, T5_TILES = 8;
constexpr int T5_BOX_W = 24, T5_BOX_H = 6, T5_BOX_X0 = 4;      // patch origin = (16*bx - 4, 4*by - 1)
constexpr int T5_K = 104, T5_STAGES = 7, T5_STAGE_Q = 16;
constexpr int T5_KCORES = T5_K / 4;                            // 26 16-byte K chunks
constexpr int T5_N = 32;
constexpr int T5_QSTRIDE = T5_BOX_W * T5_BOX_H;                // 144 floats
constexpr int T5_PATCH_FLOATS = T5_QSTRIDE * T5_K;             // 14976
constexpr int T5_PATCH_BYTES = T5_PATCH_FLOATS * 4;            // 59904
constexpr int T5_B_FLOATS = T5_KCORES * T5_N * 4;              // 3328 per operand (hi / lo)
constexpr int T5_TMEM_COLS = 256;
constexpr int T5_COL_AHI = 0, T5_COL_ALO = T5_K, T5_COL_D = 2 * T5_K;   // 0, 104, 208
constexpr int T5_THREADS = 288;
constexpr int T5_PRODUCERS = 256;
constexpr size_t T5_SMEM = (size_t)T5_PATCH_BYTES + 2 * T5_B_FLOATS * 4 + 128 * 8 + 32 * 8 + 16 + 128;

// instruction descriptor for this kernel: M = 128, N = 32 (tc5_common.cuh)
constexpr uint32_t T5_IDESC = tc5_idesc_tf32(128, T5_N);

// class probabilities -> pre-split B operand in the UMMA K-major core-matrix layout:
// p_hi / p_lo [B][26 k-chunks][32 classes][4 queries]  (element (q, c) at (q/4)*128 + c*4 + q%4)
__global__ void m2f_class_probs_umma_kernel(const float *__restrict__ cls, int B, int Q, int C1,
                                            float *__restrict__ p_hi, float *__restrict__ p_lo) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;      // over B * 104
    if (r >= B * T5_K) return;
    const int b = r / T5_K, q = r - b * T5_K;
    float *ph = p_hi + (long long)b * T5_B_FLOATS + (q >> 2) * (T5_N * 4) + (q & 3);
    float *pl = p_lo + (long long)b * T5_B_FLOATS + (q >> 2) * (T5_N * 4) + (q & 3);
    if (q >= Q) {
        for (int c = 0; c < T5_N; c++) { ph[c * 4] = 0.f; pl[c * 4] = 0.f; }
        return;
    }
    const float *x = cls + ((long long)b * Q + q) * C1;
    float m = -INFINITY;
    for (int c = 0; c < C1; c++) m = fmaxf(m, x[c]);
    float s = 0.f;
    for (int c = 0; c < C1; c++) s += expf(x[c] - m);
    for (int c = 0; c < T5_N; c++) {
        const float p = (c < C1 - 1) ? expf(x[c] - m) / s : 0.f;
        const float hi = __uint_as_float(__float_as_uint(p) & 0xFFFFE000u);
        ph[c * 4] = hi;
        pl[c * 4] = __uint_as_float(__float_as_uint(p - hi) & 0xFFFFE000u);
    }
}

template <bool HAS_EXTRA>
__global__ void __launch_bounds__(T5_THREADS, 2)
m2f_tc5_x4_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ p_hi,
                  const float *__restrict__ p_lo, int Q, int h, int w, M2FOut out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_patch = reinterpret_cast<float *>(smem_raw);                      // [104][6][24]
    float *s_bhi = s_patch + T5_PATCH_FLOATS;                                  // [26][32][4]
    float *s_blo = s_bhi + T5_B_FLOATS;
    int *s_keep = reinterpret_cast<int *>(s_blo + T5_B_FLOATS);                // [128]
    float *s_kscore = reinterpret_cast<float *>(s_keep + 128);                 // [128]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_kscore + 128);            // full[7] empty[7] d_full d_empty patch
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 32);
    uint64_t *bar_full = s_bar, *bar_empty = s_bar + T5_STAGES, *bar_dfull = s_bar + 2 * T5_STAGES,
             *bar_dempty = bar_dfull + 1, *bar_patch = bar_dfull + 2;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int sx0 = blockIdx.x * (T5_TILE_W / 4) - T5_BOX_X0, sy0 = blockIdx.y * (T5_BLOCK_H / 4) - 1;
    const int y_block = blockIdx.y * T5_BLOCK_H;
    const int n_tiles = min(T5_TILES, (out.Hc - y_block + 1) / 2);             // host guarantees y_block < Hc

    if (tid == T5_PRODUCERS) {
        for (int s = 0; s < T5_STAGES; s++) { mbar_init(&bar_full[s], T5_PRODUCERS); mbar_init(&bar_empty[s], 1); }
        mbar_init(bar_dfull, 1);
        mbar_init(bar_dempty, 128);
        mbar_init(bar_patch, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "n"(T5_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *s_tmem;

    if (tid == T5_PRODUCERS) {
        mbar_expect_tx(bar_patch, T5_PATCH_BYTES);
        tma_load_3d(s_patch, &tmap, bar_patch, sx0, sy0, b * Q);
    }
    // B operand (class-probability table, already in core-matrix order) + keep table while the patch is in flight
    {
        const float4 *gh = reinterpret_cast<const float4 *>(p_hi + (long long)b * T5_B_FLOATS);
        const float4 *gl = reinterpret_cast<const float4 *>(p_lo + (long long)b * T5_B_FLOATS);
        for (int i = tid; i < T5_B_FLOATS / 4; i += T5_THREADS) {
            reinterpret_cast<float4 *>(s_bhi)[i] = __ldg(gh + i);
            reinterpret_cast<float4 *>(s_blo)[i] = __ldg(gl + i);
        }
        if (HAS_EXTRA)
            for (int i = tid; i < 128; i += T5_THREADS) {
                s_keep[i] = (i < Q) ? out.keep_slot[(long long)b * Q + i] : -1;
                s_kscore[i] = (i < Q) ? out.keep_score[(long long)b * Q + i] : 0.f;
            }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // tensor core reads s_bhi / s_blo
    }
    mbar_wait(bar_patch, 0);
    // replicate the image edge into the zero-filled halo (torch clamps source indices)
    if ((sx0 < 0) || (sy0 < 0) || (sx0 + T5_BOX_W > w) || (sy0 + T5_BOX_H > h)) {
        for (int i = tid; i < T5_PATCH_FLOATS; i += T5_THREADS) {
            const int q = i / T5_QSTRIDE, rc = i - q * T5_QSTRIDE, r = rc / T5_BOX_W, c = rc - r * T5_BOX_W;
            const int rs = min(max(sy0 + r, 0), h - 1) - sy0, cs = min(max(sx0 + c, 0), w - 1) - sx0;
            if ((rs != r || cs != c) && rs >= 0 && rs < T5_BOX_H && cs >= 0 && cs < T5_BOX_W)
                s_patch[i] = s_patch[q * T5_QSTRIDE + rs * T5_BOX_W + cs];   // source cell is in-bounds, never rewritten
        }
    }
    __syncthreads();

    if (warp < 8) {
        // ===== producers (+ epilogue on warps 0-3) =====
        const int quarter = warp & 3, half = warp >> 2;
        const int m = quarter * 32 + lane;               // M row == TMEM lane
        const int xx = m & 63, rr = m >> 6;
        const int x = blockIdx.x * T5_TILE_W + xx;
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
        const long long plane = (long long)out.Hc * out.Wc;
        float wx1;
        {
            const float sx = 0.25f * ((float)(xx & 3) + 0.5f) - 0.5f;     // 0.625, 0.875, 0.125, 0.375
            wx1 = sx - floorf(sx);
        }
        const float wx0 = 1.f - wx1;
        const int c_off = ((xx - 2) >> 2) + T5_BOX_X0;                    // left tap column in the patch

        auto epilogue = [&](int t) {
            mbar_wait(bar_dfull, t & 1);
            tc5_fence_after();
            uint32_t d0[16], d1[4];
            tc5_ld16(lane_base + T5_COL_D, d0);
            tc5_ld4(lane_base + T5_COL_D + 16, d1);
            tc5_wait_ld();
            tc5_fence_before();
            mbar_arrive(bar_dempty);                                      // accumulator may be overwritten
            const int y = y_block + 2 * t + rr;
            if (x < out.Wc && y < out.Hc) {
                float v[M2F_C];
#pragma unroll
                for (int c = 0; c < 16; c++) v[c] = __uint_as_float(d0[c]);
#pragma unroll
                for (int c = 16; c < M2F_C; c++) v[c] = __uint_as_float(d1[c - 16]);
                const long long o = (long long)y * out.Wc + x;
                if (out.semseg) {
                    float *base = out.semseg + (long long)b * out.semseg_bstride + o;
#pragma unroll
                    for (int c = 0; c < M2F_C; c++) stg_stream_f1(base + c * plane, v[c]);
                }
                if (out.anomaly) {
                    float mx = v[0];
#pragma unroll
                    for (int c = 1; c < M2F_C; c++) mx = fmaxf(mx, v[c]);
                    stg_stream_f1(out.anomaly + (long long)b * plane + o, 1.0f - mx);
                }
            }
        };

        for (int t = 0; t < n_tiles; t++) {
            const int yy = 2 * t + rr;                                    // row inside the CTA block
            float wy1;
            {
                const float sy = 0.25f * ((float)(yy & 3) + 0.5f) - 0.5f;
                wy1 = sy - floorf(sy);
            }
            const float wy0 = 1.f - wy1;
            constexpr float NL2E = -1.4426950408889634f;
            const float w00 = NL2E * (wy0 * wx0), w01 = NL2E * (wy0 * wx1), w10 = NL2E * (wy1 * wx0), w11 = NL2E * (wy1 * wx1);
            const int r_off = ((yy - 2) >> 2) + 1;                        // upper tap row in the patch
            const float *tap = s_patch + r_off * T5_BOX_W + c_off;
            const int y = y_block + yy;
            const bool px_ok = HAS_EXTRA && x < out.Wc && y < out.Hc;

            for (int s = 0; s < T5_STAGES; s++) {
                const int q0 = s * T5_STAGE_Q + half * 8;
                if (t > 0) mbar_wait(&bar_empty[s], (t - 1) & 1);         // MMAs of the previous tile read this stage
                tc5_fence_after();
                if (q0 < T5_K) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float *p = tap + (q0 + j) * T5_QSTRIDE;
                        const float e = w00 * p[0] + w01 * p[1] + w10 * p[T5_BOX_W] + w11 * p[T5_BOX_W + 1];
                        float ex, sg;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(e));
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.0f + ex));
                        hi[j] = __float_as_uint(sg) & 0xFFFFE000u;
                        lo[j] = __float_as_uint(sg - __uint_as_float(hi[j]));
                        if (HAS_EXTRA) {
                            const int slot = s_keep[q0 + j];
                            if (slot >= 0 && px_ok)
                                out.extra[(long long)b * out.extra_bstride + (long long)slot * plane + (long long)y * out.Wc + x] =
                                    s_kscore[q0 + j] * sg;
                        }
                    }
                    tc5_st8(lane_base + T5_COL_AHI + q0, hi);
                    tc5_st8(lane_base + T5_COL_ALO + q0, lo);
                    tc5_wait_st();
                }
                tc5_fence_before();
                mbar_arrive(&bar_full[s]);
                if (s == 0 && t > 0 && warp < 4) epilogue(t - 1);         // one tile behind: the MMAs are long done
            }
        }
        if (warp < 4) epilogue(n_tiles - 1);
    } else {
        // ===== MMA issuer: the whole warp waits (stays converged), lane 0 issues =====
        const uint32_t bhi = smem_u32(s_bhi), blo = smem_u32(s_blo);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);           // warp-uniform copy for the MMA operands
        for (int t = 0; t < n_tiles; t++) {
            for (int s = 0; s < T5_STAGES; s++) {
                mbar_wait(&bar_full[s], t & 1);
                if (s == 0 && t > 0) mbar_wait(bar_dempty, (t - 1) & 1);  // epilogue has read the previous tile
                tc5_fence_after();
                if (elect_one_sync()) {
                    const int nks = (s == T5_STAGES - 1) ? 1 : 2;         // 104 = 6 x 16 + 8
                    for (int kk = 0; kk < nks; kk++) {
                        const int ks = 2 * s + kk;
                        const uint64_t dh = tc5_smem_desc(bhi + ks * 2 * (T5_N * 16), T5_N * 16, 128);
                        const uint64_t dl = tc5_smem_desc(blo + ks * 2 * (T5_N * 16), T5_N * 16, 128);
                        tc5_mma_ts(tmem_u + T5_COL_D, tmem_u + T5_COL_ALO + ks * 8, dh, T5_IDESC, ks > 0);
                        tc5_mma_ts(tmem_u + T5_COL_D, tmem_u + T5_COL_AHI + ks * 8, dl, T5_IDESC, 1);
                        tc5_mma_ts(tmem_u + T5_COL_D, tmem_u + T5_COL_AHI + ks * 8, dh, T5_IDESC, 1);
                    }
                    tc5_commit(&bar_empty[s]);
                    if (s == T5_STAGES - 1) tc5_commit(bar_dfull);
                }
                __syncwarp();
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc5_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(T5_TMEM_COLS) : "memory");
    }
}

}  // namespace mss
