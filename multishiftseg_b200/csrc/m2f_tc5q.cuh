// Mask2Former fused post-head inference, tcgen05 variant with SHARED TAPS (exact x4 upsample path) -- the
// default fast path.  Same contract as the other x4 kernels.
//
// (Round-1 timing experiments -- dropping either MUFU op of the sigmoid 146 -> 118 us/image, one TMEM store instead of
// four: no change, a quarter of the MMAs: -17 % before the elect_one fix, none after -- were run with compile-time
// switches that have since been removed from this file; their results are in DESIGN.md section 4.)
//
// ncu on the first tcgen05 kernel (m2f_tc5.cuh, round 1, 204 us/image): 22.9 warp instructions per
// (pixel, query): 4 LDS + 8 FMA-pipe ops of bilinear interpolation, 5 of sigmoid + TF32 split, ~5 of address /
// loop / barrier overhead; issue slots 68 % busy, MUFU pipe 50 %.  With x4 upsampling (align_corners=False)
// the four output pixels x = 4k+2 .. 4k+5 of one row all interpolate between the SAME two source columns
// (k, k+1) and the same two source rows, so here one thread owns such a 4 x 1 pixel block:
//     4 LDS + 10 FMA-pipe ops per (block, query)  =  3.5 instead of 12 per (pixel, query),
// which leaves the two MUFU ops of the sigmoid (ex2 + rcp, 16 lanes/clk/SM) as the binding pipe.
//
// A thread <-> one TMEM lane, so its 4 pixels go to 4 DIFFERENT M-tiles that are in flight together:
//   M-tile i (i = 0..3) = pixel i of each of 128 blocks;  a "group" = 128 blocks = 8 rows x 64 px.
//   CTA = 64 x 32 output pixels (x from 64*bx - 2: blocks start at x = 2 mod 4) = 4 groups, one TMA box
//   (20 x 10 x 104 fp32, 83.2 KB) with the low-res patch of ALL queries; out-of-image halo cells are
//   overwritten with the edge value (torch clamps source indices), after which every tap is border-free.
//   The contraction  semseg[px, c] = sum_q S[px, q] P[q, c]  is a split-TF32 GEMM with fp32 accumulation in
//   TMEM: S = S_hi + S_lo, P = P_hi + P_lo (hi = top 19 bits, lo = exact remainder).  One K = 8 instruction
//   covers FOUR queries: its A columns are [S_hi(q..q+3) | S_lo(q..q+3)] -- exactly what one thread produces,
//   one 8-column tcgen05.st -- and it is issued twice, against B = [P_hi ; P_hi] and B = [P_lo ; P_lo]
//   (the same 4-query core matrix for both K chunks: descriptor LBO = 0, or a duplicated table), which sums
//   all four partial products (S_hi + S_lo)(P_hi + P_lo).
//   * warps 0-7 (producers): warp = (lane quarter, half); per stage of 8 queries a thread evaluates 4 queries
//     (its half) x 4 pixels and writes them straight into TENSOR MEMORY as the A operands of the 4 tiles; two
//     A buffers alternate;
//   * warp 8, one lane: per stage 4 tiles x 4 tcgen05.mma (M = 128, N = 32, K = 8, A from TMEM, B = class-
//     probability table in shared memory), tcgen05.commit frees the A buffer / publishes the accumulators;
//   * a producer evaluates its stage BEFORE waiting for the A buffer to be free, so the MMAs of two stages
//     ago have a whole stage of slack;
//   * epilogue (all producer warps, one stage into the next group so the last MMAs are long done): tcgen05.ld of the accumulators; half h of a quarter
//     takes pixels 2h, 2h+1 of each block, so every class is one 8-byte store per thread and a warp writes
//     contiguous 256-byte runs; 1 - max_c for the anomaly map.
// TMEM: D tiles at columns 0..127, A buffers at 128..255 (per buffer: tile i, half h at i*16 + 8*h)
//   -> 256 columns per CTA, two CTAs per SM.
#pragma once

namespace mss {

constexpr int TQ_W = 64, TQ_H = 32, TQ_GROUP_H = 8, TQ_GROUPS = 4;
constexpr int TQ_BOX_W = 20, TQ_BOX_H = TQ_H / 4 + 2, TQ_BOX_X0 = 4;   // patch origin = (16*bx - 4, (TQ_H/4)*by - 1)
constexpr int TQ_KSTEPS = T5_K / 8;                            // 13
constexpr int TQ_QSTRIDE = TQ_BOX_W * TQ_BOX_H;                // 200 floats
constexpr int TQ_PATCH_FLOATS = TQ_QSTRIDE * T5_K;             // 20800
constexpr int TQ_PATCH_BYTES = TQ_PATCH_FLOATS * 4;            // 83200
constexpr int TQ_TMEM_COLS = 256;
constexpr int TQ_COL_D = 0, TQ_COL_A = 128;                    // D tile i: i*32; A buffer s: 128 + 64*s
constexpr int TQ_THREADS = 288, TQ_PRODUCERS = 256;
constexpr size_t tq_smem(bool dup) {
    return (size_t)TQ_PATCH_BYTES + (dup ? 4 : 2) * T5_B_FLOATS * 4 + 128 * 8 + 16 * 8 + 16 + 128;
}

__device__ __forceinline__ void stg_stream_f2(float *p, float a, float b) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// DUP_B: keep two copies of every 4-query core matrix of the class table in shared memory (K chunk 0 and 1)
// instead of pointing both K chunks of the descriptor at the same one (LBO = 0)
// 1 / d for d in [1, 2^127) on the FMA pipe: integer-subtract seed (|1 - d y0| <= 0.0506), one cubic and one
// quadratic Newton step (error 0.0506^3 -> squared = 1.7e-8, below float32 rounding).  Used for TQ_RCP_FMA of
// the 4 pixels of a block to take load off the MUFU pipe, which binds this kernel (timing experiments, round 1:
// dropping either MUFU of the sigmoid took 146 -> 118 us/image).
#ifndef TQ_RCP_FMA
#define TQ_RCP_FMA 0     /* measured: 0 -> 145.1, 1 -> 148.2, 2 -> 147.2, 3 -> 151.5 us/image (issue slots co-limit) */
#endif
// TQ_RCP_PAIR: one MUFU.RCP for two sigmoids -- r = rcp(a0 * a1), 1/a0 = r * a1, 1/a1 = r * a0 (a = 1 + 2^e, e clamped
// to 63 so the product stays finite) -- 1.5 instead of 2 MUFU per sigmoid at the price of 1.5 more FMA/ALU instructions.
#ifndef TQ_RCP_PAIR
#define TQ_RCP_PAIR 0     /* measured: 0 -> 143.6, 1 -> 143.3 us/image (neutral: the MUFU pipe is not alone on the critical path) */
#endif
#ifndef TQ_MMA_SLEEP_NS
#define TQ_MMA_SLEEP_NS 32
#endif
__device__ __forceinline__ float rcp_fma(float d) {
    const float y0 = __int_as_float(0x7EF31000 - __float_as_int(d));
    const float r = fmaf(-d, y0, 1.0f);
    const float y1 = fmaf(y0, fmaf(r, r, r), y0);
    return fmaf(y1, fmaf(-d, y1, 1.0f), y1);
}

template <bool HAS_EXTRA, bool DUP_B>
__global__ void __launch_bounds__(TQ_THREADS, 2)
m2f_tc5q_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ p_hi,
                const float *__restrict__ p_lo, int Q, int h, int w, M2FOut out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_patch = reinterpret_cast<float *>(smem_raw);                      // [104][10][20]
    constexpr int BF = (DUP_B ? 2 : 1) * T5_B_FLOATS;
    float *s_bhi = s_patch + TQ_PATCH_FLOATS;                                  // [26][1 or 2][32][4]
    float *s_blo = s_bhi + BF;
    int *s_keep = reinterpret_cast<int *>(s_blo + BF);                         // [128]
    float *s_kscore = reinterpret_cast<float *>(s_keep + 128);                 // [128]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_kscore + 128);            // full[2] empty[2] d_full d_empty patch
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 16);
    uint64_t *bar_full = s_bar, *bar_empty = s_bar + 2, *bar_dfull = s_bar + 4, *bar_dempty = s_bar + 5,
             *bar_patch = s_bar + 6;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int sx0 = blockIdx.x * (TQ_W / 4) - TQ_BOX_X0, sy0 = blockIdx.y * (TQ_H / 4) - 1;
    const int y_block = blockIdx.y * TQ_H;
    const int n_groups = min(TQ_GROUPS, (out.Hc - y_block + TQ_GROUP_H - 1) / TQ_GROUP_H);   // host: y_block < Hc

    if (tid == TQ_PRODUCERS) {
        for (int s = 0; s < 2; s++) { mbar_init(&bar_full[s], TQ_PRODUCERS); mbar_init(&bar_empty[s], 1); }
        mbar_init(bar_dfull, 1);
        mbar_init(bar_dempty, TQ_PRODUCERS);
        mbar_init(bar_patch, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "n"(TQ_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *s_tmem;

    if (tid == TQ_PRODUCERS) {
        mbar_expect_tx(bar_patch, TQ_PATCH_BYTES);
        tma_load_3d(s_patch, &tmap, bar_patch, sx0, sy0, b * Q);
    }
    // B operand (class-probability table, already in core-matrix order) + keep table while the patch is in flight
    {
        const float4 *gh = reinterpret_cast<const float4 *>(p_hi + (long long)b * T5_B_FLOATS);
        const float4 *gl = reinterpret_cast<const float4 *>(p_lo + (long long)b * T5_B_FLOATS);
        for (int i = tid; i < T5_B_FLOATS / 4; i += TQ_THREADS) {
            const float4 vh = __ldg(gh + i), vl = __ldg(gl + i);
            if (DUP_B) {
                const int o = (i >> 5) * 64 + (i & 31);                        // chunk k' -> [k'][0][c], [k'][1][c]
                reinterpret_cast<float4 *>(s_bhi)[o] = vh; reinterpret_cast<float4 *>(s_bhi)[o + 32] = vh;
                reinterpret_cast<float4 *>(s_blo)[o] = vl; reinterpret_cast<float4 *>(s_blo)[o + 32] = vl;
            } else {
                reinterpret_cast<float4 *>(s_bhi)[i] = vh;
                reinterpret_cast<float4 *>(s_blo)[i] = vl;
            }
        }
        if (HAS_EXTRA)
            for (int i = tid; i < 128; i += TQ_THREADS) {
                s_keep[i] = (i < Q) ? out.keep_slot[(long long)b * Q + i] : -1;
                s_kscore[i] = (i < Q) ? out.keep_score[(long long)b * Q + i] : 0.f;
            }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // tensor core reads s_bhi / s_blo
    }
    mbar_wait(bar_patch, 0);
    // replicate the image edge into the zero-filled halo (torch clamps source indices)
    if ((sx0 < 0) || (sy0 < 0) || (sx0 + TQ_BOX_W > w) || (sy0 + TQ_BOX_H > h)) {
        for (int i = tid; i < TQ_PATCH_FLOATS; i += TQ_THREADS) {
            const int q = i / TQ_QSTRIDE, rc = i - q * TQ_QSTRIDE, r = rc / TQ_BOX_W, c = rc - r * TQ_BOX_W;
            const int rs = min(max(sy0 + r, 0), h - 1) - sy0, cs = min(max(sx0 + c, 0), w - 1) - sx0;
            if ((rs != r || cs != c) && rs >= 0 && rs < TQ_BOX_H && cs >= 0 && cs < TQ_BOX_W)
                s_patch[i] = s_patch[q * TQ_QSTRIDE + rs * TQ_BOX_W + cs];   // source cell is in-bounds, never rewritten
        }
    }
    __syncthreads();

    if (warp < 8) {
        // ===== producers + epilogue =====
        const int quarter = warp & 3, half = warp >> 2;
        const int m = quarter * 32 + lane;               // block index inside the group == TMEM lane
        const int jb = m & 15, rowg = m >> 4;            // block column (0..15), row inside the group (0..7)
        const int xb = blockIdx.x * TQ_W - 2 + 4 * jb;   // x of pixel 0 of the block (== 2 mod 4)
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
        const long long plane = (long long)out.Hc * out.Wc;
        constexpr float NL2E = -1.4426950408889634f;
        // pixel i of a block sits at source coordinate k + 0.125 + 0.25 i between columns k and k + 1
        const float wxs0 = NL2E * 0.125f, wxs1 = NL2E * 0.375f, wxs2 = NL2E * 0.625f, wxs3 = NL2E * 0.875f;
        const float *tap_col = s_patch + (jb + TQ_BOX_X0 - 1);      // column k = 16*bx + jb - 1 of the patch
        const bool vec2 = ((out.Wc & 1) == 0);

        // epilogue of group g: half h stores pixels 2h, 2h+1 of every block
        auto epilogue = [&](int g) {
            const int y = y_block + g * TQ_GROUP_H + rowg;
            mbar_wait(bar_dfull, g & 1);
            tc5_fence_after();
            const int x0 = xb + 2 * half;
            const bool ok0 = x0 >= 0 && x0 < out.Wc && y < out.Hc, ok1 = x0 + 1 >= 0 && x0 + 1 < out.Wc && y < out.Hc;
            const long long o = (long long)y * out.Wc + x0;
            float *sbase = out.semseg ? out.semseg + (long long)b * out.semseg_bstride + o : nullptr;
            const bool sv2 = vec2 && ok0 && ok1 && ((out.semseg_bstride & 1) == 0) && ((((uintptr_t)out.semseg) & 7) == 0);
            float mx0 = -INFINITY, mx1 = -INFINITY;
            const uint32_t d_a = lane_base + TQ_COL_D + (2 * half) * 32, d_b = d_a + 32;
#pragma unroll
            for (int c0 = 0; c0 < 24; c0 += 8) {
                uint32_t va[8], vb[8];
                tc5_ld8(d_a + c0, va);
                tc5_ld8(d_b + c0, vb);
                tc5_wait_ld();
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    if (c0 + c < M2F_C) {
                        const float fa = __uint_as_float(va[c]), fb = __uint_as_float(vb[c]);
                        mx0 = fmaxf(mx0, fa);
                        mx1 = fmaxf(mx1, fb);
                        if (sbase) {
                            float *pc = sbase + (long long)(c0 + c) * plane;
                            if (sv2) stg_stream_f2(pc, fa, fb);
                            else {
                                if (ok0) stg_stream_f1(pc, fa);
                                if (ok1) stg_stream_f1(pc + 1, fb);
                            }
                        }
                    }
                }
            }
            tc5_fence_before();
            mbar_arrive(bar_dempty);                                // accumulators may be overwritten
            if (out.anomaly) {
                float *pa = out.anomaly + (long long)b * plane + o;
                if (vec2 && ok0 && ok1 && ((((uintptr_t)out.anomaly) & 7) == 0)) stg_stream_f2(pa, 1.0f - mx0, 1.0f - mx1);
                else {
                    if (ok0) stg_stream_f1(pa, 1.0f - mx0);
                    if (ok1) stg_stream_f1(pa + 1, 1.0f - mx1);
                }
            }
        };

        for (int g = 0; g < n_groups; g++) {
            const int yy = g * TQ_GROUP_H + rowg;                   // row inside the CTA block
            const int y = y_block + yy;
            float wy1;
            {
                const float sy = 0.25f * ((float)(yy & 3) + 0.5f) - 0.5f;
                wy1 = sy - floorf(sy);
            }
            const int r_off = ((yy - 2) >> 2) + 1;                  // upper tap row in the patch
            const float *tap = tap_col + r_off * TQ_BOX_W + half * 4 * TQ_QSTRIDE;

            for (int ks = 0; ks < TQ_KSTEPS; ks++, tap += 8 * TQ_QSTRIDE) {
                const int u = g * TQ_KSTEPS + ks, slot = u & 1;
                const int q0 = ks * 8 + half * 4;
                const uint32_t a_base = lane_base + TQ_COL_A + slot * 64 + half * 8;
                if (q0 < Q) {                                       // Q % 4 == 0 on this path (host-checked)
                    uint32_t v[4][8];                               // [pixel / tile][hi q0..q0+3 | lo q0..q0+3]
                    // Software pipeline over the 4 queries: the ex2 of query j are issued together with the rcp of
                    // query j-1 (volatile asm keeps that order), so a warp offers the MUFU pipe -- the binding pipe,
                    // 8 issue cycles per warp instruction -- a steady trickle instead of bursts of 16 + 16.
                    float ex[4];
#pragma unroll
                    for (int j = 0; j <= 4; j++) {
                        float e[4];
                        if (j < 4) {
                            const float *p = tap + j * TQ_QSTRIDE;
                            const float a = p[0], bb = p[1], c = p[TQ_BOX_W], d = p[TQ_BOX_W + 1];
                            const float L = fmaf(wy1, c - a, a), R = fmaf(wy1, d - bb, bb);
                            const float D = R - L, Ls = NL2E * L;
                            e[0] = fmaf(wxs0, D, Ls); e[1] = fmaf(wxs1, D, Ls); e[2] = fmaf(wxs2, D, Ls); e[3] = fmaf(wxs3, D, Ls);
                        }
#if TQ_RCP_PAIR
                        float sgp[4];
                        if (j > 0) {
#pragma unroll
                            for (int i = 0; i < 4; i += 2) {
                                const float a0 = 1.0f + ex[i], a1 = 1.0f + ex[i + 1];
                                float r;
                                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a0 * a1));
                                sgp[i] = r * a1;
                                sgp[i + 1] = r * a0;
                            }
                        }
#endif
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            float sg = 0.f;
#if TQ_RCP_PAIR
                            if (j > 0) sg = sgp[i];
                            if (false) {
#else
                            if (j > 0) {
#endif
                                if (i < TQ_RCP_FMA) sg = rcp_fma(1.0f + ex[i]);
                                else asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.0f + ex[i]));
                            }
                            if (j < 4) {
                                // (the FMA-pipe reciprocal needs a finite 1 + 2^e: clamp e, i.e. mask logits below -87)
                                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex[i]) : "f"(TQ_RCP_PAIR ? fminf(e[i], 63.f) : (i < TQ_RCP_FMA ? fminf(e[i], 126.f) : e[i])));
                            }
                            if (j > 0) {
                                const int jq = j - 1;
                                v[i][jq] = __float_as_uint(sg) & 0xFFFFE000u;
                                v[i][4 + jq] = __float_as_uint(sg - __uint_as_float(v[i][jq]));
                                if (HAS_EXTRA) {
                                    const int slot_k = s_keep[q0 + jq];
                                    const int x = xb + i;
                                    if (slot_k >= 0 && x >= 0 && x < out.Wc && y < out.Hc)
                                        out.extra[(long long)b * out.extra_bstride + (long long)slot_k * plane + (long long)y * out.Wc + x] =
                                            s_kscore[q0 + jq] * sg;
                                }
                            }
                        }
                    }
                    // the values are ready: only now wait for the MMAs of use u-2 to have read this A buffer
                    if (u >= 2) mbar_wait(&bar_empty[slot], ((u >> 1) + 1) & 1);
                    tc5_fence_after();
#pragma unroll
                    for (int i = 0; i < 4; i++) tc5_st8(a_base + i * 16, v[i]);
                } else {
                    // padded queries: the box holds the next image's masks there; they must not reach the MMA
                    if (u >= 2) mbar_wait(&bar_empty[slot], ((u >> 1) + 1) & 1);
                    tc5_fence_after();
                    const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
                    for (int i = 0; i < 4; i++) tc5_st8(a_base + i * 16, z);
                }
                tc5_wait_st();
                tc5_fence_before();
                mbar_arrive(&bar_full[slot]);
                // epilogue of the previous group, one stage late: its last MMAs have had a whole stage to finish
                if (ks == 0 && g > 0) epilogue(g - 1);
            }
        }
        epilogue(n_groups - 1);
    } else {
        // ===== MMA issuer: the whole warp waits (stays converged), lane 0 issues =====
        const uint32_t bhi = smem_u32(s_bhi), blo = smem_u32(s_blo);
        // warp-uniform copy of the TMEM base + elect_one_sync() below: tcgen05.mma takes its operands from uniform
        // registers; with a per-thread base under `if (lane == 0)` every MMA sat in an ELECT / R2UR.BROADCAST /
        // BRA.U.ANY waterfall loop (~100 cycles per MMA), which made MMA issue the critical path of the kernel
        // (ncu + timing experiments, round 1)
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        for (int g = 0; g < n_groups; g++) {
            for (int ks = 0; ks < TQ_KSTEPS; ks++) {
                const int u = g * TQ_KSTEPS + ks, slot = u & 1;
                mbar_wait_backoff(&bar_full[slot], (u >> 1) & 1, TQ_MMA_SLEEP_NS);
                if (ks == 0 && g > 0) mbar_wait_backoff(bar_dempty, (g - 1) & 1, TQ_MMA_SLEEP_NS);   // epilogue has read the previous group
                tc5_fence_after();
                if (elect_one_sync()) {
                    // 4-query core matrix k' = 2 ks + hh of the class table, used for both K chunks
                    constexpr uint32_t CHUNK = T5_N * 16 * (DUP_B ? 2 : 1), LBO = DUP_B ? T5_N * 16 : 0;
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const uint64_t dh = tc5_smem_desc(bhi + (2 * ks + hh) * CHUNK, LBO, 128);
                        const uint64_t dl = tc5_smem_desc(blo + (2 * ks + hh) * CHUNK, LBO, 128);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint32_t d = tmem_u + TQ_COL_D + i * 32, a = tmem_u + TQ_COL_A + slot * 64 + i * 16 + hh * 8;
                            tc5_mma_ts(d, a, dh, T5_IDESC, (ks | hh) > 0);
                            tc5_mma_ts(d, a, dl, T5_IDESC, 1);
                        }
                    }
                    tc5_commit(&bar_empty[slot]);
                    if (ks == TQ_KSTEPS - 1) tc5_commit(bar_dfull);
                }
                __syncwarp();
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc5_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TQ_TMEM_COLS) : "memory");
    }
}

}  // namespace mss
