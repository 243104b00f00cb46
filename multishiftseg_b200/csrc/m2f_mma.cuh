// Mask2Former fused post-head inference, tensor-core variant (exact x4 upsample path).
//
// Same contract as m2f_fused_x4_kernel, but the 100 x 19 contraction  semseg[px, c] = sum_q S[px, q] P[q, c]
// runs on the tensor cores as a 3xTF32 split GEMM (mma.sync.m16n8k8, fp32 accumulate):
//     S = S_hi + S_lo,  P = P_hi + P_lo  (hi = top 19 bits, lo = exact remainder)
//     D += S_lo*P_hi;  D += S_hi*P_lo;  D += S_hi*P_hi          (S_lo*P_lo ~ 2^-22, dropped)
// which keeps fp32-level accuracy (the 1e-5 parity bar) while the CUDA cores only interpolate, evaluate
// the sigmoid and split.  ncu on the FFMA variant (round 1): fma pipe 54 %, issue 66 %, 8 warps/SM -- the
// 152 FFMA per query per thread were the issue-slot hog; here they become 9 MMA per 16 px x 8 queries and
// the kernel is bounded by the MUFU pipe (2 per sigmoid), not by the FMA pipe.
//
// Geometry: CTA = 4 warps, output tile 64 x 8 px; warp = 2 output rows x 64 cols = 8 m-tiles of 16 px
// (m-tile row r <-> pixel (y + r/8, x0 + 8*mt + r%8), so the two pixels a thread feeds into one A fragment
// are vertically adjacent and share their four source taps).  K = queries (13 k-steps of 8, Q padded to
// 104), N = classes (3 n-tiles of 8, C padded to 24).  The low-res patch of ALL queries for the tile
// (24 x 5 x 104 fp32, 48.75 KB; 5 rows instead of 4 so that the query stride is 120 = 24 mod 32 words and
// the four query-lanes of a fragment hit different banks) arrives in ONE TMA box; three CTAs per SM overlap
// each other's loads.  Image borders: the zero-filled out-of-bounds halo is overwritten in smem with the
// edge value (torch clamps indices), after which every thread runs the same border-free code.
#pragma once

namespace mss {

constexpr int MM_TILE_W = 64, MM_TILE_H = 8;
constexpr int MM_BOX_W = 24, MM_BOX_H = 5, MM_BOX_X0 = 4;   // patch origin = (16*bx - 4, 2*by - 1)
constexpr int MM_QPAD = 104, MM_KSTEPS = 13, MM_NPAD = 24;
constexpr int MM_QSTRIDE = MM_BOX_W * MM_BOX_H;             // 120 floats
constexpr int MM_PATCH_FLOATS = MM_QSTRIDE * MM_QPAD;       // 12480
constexpr int MM_PATCH_BYTES = MM_PATCH_FLOATS * 4;         // 49920

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// class probabilities -> padded, pre-split B operand: p_hi/p_lo [B][104][24]
__global__ void m2f_class_probs_split_kernel(const float *__restrict__ cls, int B, int Q, int C1,
                                             float *__restrict__ p_hi, float *__restrict__ p_lo) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;      // over B * 104
    if (r >= B * MM_QPAD) return;
    const int b = r / MM_QPAD, q = r - b * MM_QPAD;
    float *ph = p_hi + (long long)r * MM_NPAD, *pl = p_lo + (long long)r * MM_NPAD;
    if (q >= Q) {
        for (int c = 0; c < MM_NPAD; c++) { ph[c] = 0.f; pl[c] = 0.f; }
        return;
    }
    const float *x = cls + ((long long)b * Q + q) * C1;
    float m = -INFINITY;
    for (int c = 0; c < C1; c++) m = fmaxf(m, x[c]);
    float s = 0.f;
    for (int c = 0; c < C1; c++) s += expf(x[c] - m);
    for (int c = 0; c < MM_NPAD; c++) {
        const float p = (c < C1 - 1) ? expf(x[c] - m) / s : 0.f;
        const float hi = __uint_as_float(__float_as_uint(p) & 0xFFFFE000u);
        ph[c] = hi;
        pl[c] = __uint_as_float(__float_as_uint(p - hi) & 0xFFFFE000u);
    }
}

__global__ void __launch_bounds__(128, 3)
m2f_mma_x4_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ p_hi,
                  const float *__restrict__ p_lo, int Q, int h, int w, M2FOut out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_patch = reinterpret_cast<float *>(smem_raw);                   // [104][5][24]
    float *s_ph = s_patch + MM_PATCH_FLOATS;                                // [104][24]
    float *s_pl = s_ph + MM_QPAD * MM_NPAD;                                 // [104][24]
    int *s_keep = reinterpret_cast<int *>(s_pl + MM_QPAD * MM_NPAD);        // [104]
    float *s_kscore = reinterpret_cast<float *>(s_keep + MM_QPAD);          // [104]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_kscore + MM_QPAD);

    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.z;
    const int sx0 = blockIdx.x * (MM_TILE_W / 4) - MM_BOX_X0, sy0 = blockIdx.y * (MM_TILE_H / 4) - 1;

    if (tid == 0) {
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(s_bar, MM_PATCH_BYTES);
        tma_load_3d(s_patch, &tmap, s_bar, sx0, sy0, b * Q);
    }
    // B operand + keep table while the patch is in flight
    {
        const float4 *gh = reinterpret_cast<const float4 *>(p_hi + (long long)b * MM_QPAD * MM_NPAD);
        const float4 *gl = reinterpret_cast<const float4 *>(p_lo + (long long)b * MM_QPAD * MM_NPAD);
        for (int i = tid; i < MM_QPAD * MM_NPAD / 4; i += 128) {
            reinterpret_cast<float4 *>(s_ph)[i] = __ldg(gh + i);
            reinterpret_cast<float4 *>(s_pl)[i] = __ldg(gl + i);
        }
        const bool has_extra = out.keep_slot != nullptr && out.extra != nullptr;
        for (int i = tid; i < MM_QPAD; i += 128) {
            s_keep[i] = (has_extra && i < Q) ? out.keep_slot[(long long)b * Q + i] : -1;
            s_kscore[i] = (has_extra && i < Q) ? out.keep_score[(long long)b * Q + i] : 0.f;
        }
    }
    // per-thread geometry (border-free thanks to the halo fix-up below)
    const int x_base = blockIdx.x * MM_TILE_W + g;                 // + 8*mt
    const int y0 = blockIdx.y * MM_TILE_H + 2 * wp;                // rows y0, y0+1
    float wx1, wyA1, wyB1;
    {
        // src = 0.25*(dst+0.5)-0.5 (unclamped); l1 = src - floor(src): 0.625,0.875,0.125,0.375 for dst%4 = 0..3
        const float sx = 0.25f * ((float)(g & 3) + 0.5f) - 0.5f;
        wx1 = sx - floorf(sx);
        const float sa = 0.25f * ((float)((2 * wp) & 3) + 0.5f) - 0.5f, sb = 0.25f * ((float)((2 * wp + 1) & 3) + 0.5f) - 0.5f;
        wyA1 = sa - floorf(sa);
        wyB1 = sb - floorf(sb);
    }
    const float wx0 = 1.f - wx1, wyA0 = 1.f - wyA1, wyB0 = 1.f - wyB1;
    const int cA = MM_BOX_X0 + ((g - 2) >> 2);                     // + 2*mt ; second tap = +1
    const int rA = ((2 * wp - 2) >> 2) + 1;                        // second tap row = +1
    const int tap0 = rA * MM_BOX_W + cA;

    float acc[8][3][4];
#pragma unroll
    for (int mt = 0; mt < 8; mt++)
#pragma unroll
        for (int nt = 0; nt < 3; nt++)
#pragma unroll
            for (int i = 0; i < 4; i++) acc[mt][nt][i] = 0.f;

    mbar_wait(s_bar, 0);
    // replicate the image edge into the zero-filled halo (torch clamps source indices)
    const bool edge = (sx0 < 0) || (sy0 < 0) || (sx0 + MM_BOX_W > w) || (sy0 + MM_BOX_H > h);
    if (edge) {
        for (int i = tid; i < MM_PATCH_FLOATS; i += 128) {
            const int q = i / MM_QSTRIDE, rc = i - q * MM_QSTRIDE, r = rc / MM_BOX_W, c = rc - r * MM_BOX_W;
            const int rs = min(max(sy0 + r, 0), h - 1) - sy0, cs = min(max(sx0 + c, 0), w - 1) - sx0;
            if ((rs != r || cs != c) && rs >= 0 && rs < MM_BOX_H && cs >= 0 && cs < MM_BOX_W)
                s_patch[i] = s_patch[q * MM_QSTRIDE + rs * MM_BOX_W + cs];   // source cell is in-bounds, never rewritten
        }
    }
    __syncthreads();

    const bool row0_ok = y0 < out.Hc, row1_ok = y0 + 1 < out.Hc;
#pragma unroll 1
    for (int ks = 0; ks < MM_KSTEPS; ks++) {
        const int q0 = 8 * ks + t, q1 = q0 + 4;
        uint32_t bh[3][2], bl[3][2];
#pragma unroll
        for (int nt = 0; nt < 3; nt++) {
            bh[nt][0] = __float_as_uint(s_ph[q0 * MM_NPAD + nt * 8 + g]);
            bh[nt][1] = __float_as_uint(s_ph[q1 * MM_NPAD + nt * 8 + g]);
            bl[nt][0] = __float_as_uint(s_pl[q0 * MM_NPAD + nt * 8 + g]);
            bl[nt][1] = __float_as_uint(s_pl[q1 * MM_NPAD + nt * 8 + g]);
        }
        const float *pq0 = s_patch + q0 * MM_QSTRIDE + tap0, *pq1 = s_patch + q1 * MM_QSTRIDE + tap0;
        const bool v0 = q0 < Q, v1 = q1 < Q;
        const int slot0 = s_keep[q0], slot1 = s_keep[q1];
#pragma unroll
        for (int mt = 0; mt < 8; mt++) {
            float s[4];
            {
                const float a = pq0[2 * mt], bb = pq0[2 * mt + 1], c = pq0[2 * mt + MM_BOX_W], d = pq0[2 * mt + MM_BOX_W + 1];
                const float h0 = wx0 * a + wx1 * bb, h1 = wx0 * c + wx1 * d;
                s[0] = v0 ? sigmoid_fast(wyA0 * h0 + wyA1 * h1) : 0.f;
                s[1] = v0 ? sigmoid_fast(wyB0 * h0 + wyB1 * h1) : 0.f;
            }
            {
                const float a = pq1[2 * mt], bb = pq1[2 * mt + 1], c = pq1[2 * mt + MM_BOX_W], d = pq1[2 * mt + MM_BOX_W + 1];
                const float h0 = wx0 * a + wx1 * bb, h1 = wx0 * c + wx1 * d;
                s[2] = v1 ? sigmoid_fast(wyA0 * h0 + wyA1 * h1) : 0.f;
                s[3] = v1 ? sigmoid_fast(wyB0 * h0 + wyB1 * h1) : 0.f;
            }
            uint32_t ah[4], al[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                ah[i] = __float_as_uint(s[i]) & 0xFFFFE000u;
                al[i] = __float_as_uint(s[i] - __uint_as_float(ah[i])) & 0xFFFFE000u;
            }
#pragma unroll
            for (int nt = 0; nt < 3; nt++) {
                mma_tf32(acc[mt][nt], al, bh[nt][0], bh[nt][1]);
                mma_tf32(acc[mt][nt], ah, bl[nt][0], bl[nt][1]);
                mma_tf32(acc[mt][nt], ah, bh[nt][0], bh[nt][1]);
            }
            if ((slot0 >= 0) | (slot1 >= 0)) {     // rare: maskformer_model.py:346-352 extra channels
                const int x = x_base + 8 * mt;
                if (x < out.Wc) {
                    if (slot0 >= 0) {
                        float *e = out.extra + (long long)b * out.extra_bstride + (long long)slot0 * out.Hc * out.Wc;
                        if (row0_ok) e[(long long)y0 * out.Wc + x] = s_kscore[q0] * s[0];
                        if (row1_ok) e[(long long)(y0 + 1) * out.Wc + x] = s_kscore[q0] * s[1];
                    }
                    if (slot1 >= 0) {
                        float *e = out.extra + (long long)b * out.extra_bstride + (long long)slot1 * out.Hc * out.Wc;
                        if (row0_ok) e[(long long)y0 * out.Wc + x] = s_kscore[q1] * s[2];
                        if (row1_ok) e[(long long)(y0 + 1) * out.Wc + x] = s_kscore[q1] * s[3];
                    }
                }
            }
        }
    }

    // epilogue.  C fragment: acc[mt][nt][0..1] = (row y0,   classes 8nt+2t, +1), [2..3] = (row y0+1, same classes)
    const long long plane = (long long)out.Hc * out.Wc;
#pragma unroll
    for (int mt = 0; mt < 8; mt++) {
        const int x = x_base + 8 * mt;
        const bool xin = x < out.Wc;
        if (out.semseg && xin) {
            float *base = out.semseg + (long long)b * out.semseg_bstride + x;
#pragma unroll
            for (int nt = 0; nt < 3; nt++) {
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const int c = 8 * nt + 2 * t + i;
                    if (c < M2F_C) {
                        if (row0_ok) base[c * plane + (long long)y0 * out.Wc] = acc[mt][nt][i];
                        if (row1_ok) base[c * plane + (long long)(y0 + 1) * out.Wc] = acc[mt][nt][2 + i];
                    }
                }
            }
        }
        if (out.anomaly) {
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 3; nt++) {
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    if (8 * nt + 2 * t + i < M2F_C) {
                        m0 = fmaxf(m0, acc[mt][nt][i]);
                        m1 = fmaxf(m1, acc[mt][nt][2 + i]);
                    }
                }
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            if (xin && t == 0) {
                float *d = out.anomaly + (long long)b * plane + x;
                if (row0_ok) d[(long long)y0 * out.Wc] = 1.0f - m0;
                if (row1_ok) d[(long long)(y0 + 1) * out.Wc] = 1.0f - m1;
            }
        }
    }
}

}  // namespace mss
