// (a4-a7) Mask2Former fused post-head inference.
//
// Replaces, in ONE kernel and without materialising any [Q, Hp, Wp] tensor:
//   F.interpolate(pred_masks, size=(Hp,Wp), mode="bilinear", align_corners=False)   maskformer_model.py:264-277
//   MaskFormer.semantic_inference: softmax(cls)[..., :-1], sigmoid, einsum "qc,qhw->chw",
//       + the K extra "confident thing query" channels                                 maskformer_model.py:341-354
//   sem_seg_postprocess crop                                                           maskformer_model.py:299-300
//   TrainM2FOOD.get_anomaly_score: einsum "bqc,bqhw->bchw", crop, 1 - max_c            train_m2f.py:387-407
//
// Fast path (exact x4 upsample, the model's only configuration: common stride 4, anomaly_ft.yaml:30):
//   CTA = 256 threads = 32 x 8, output tile 128 x 16 px, thread = 2 rows x 4 cols (8 px, 152 fp32
//   accumulators).  The 6 x 34 low-res patch of every query is staged by TMA (cp.async.bulk.tensor.3d,
//   box 40 x 6 x QCHUNK (16-byte aligned start), zero fill outside the image, indices clamped at read time to reproduce
//   torch's edge replication) through a 3-stage mbarrier ring.  Per query and thread: 6 LDS, 16 lerp
//   ops, 8 sigmoids (ex2.approx + rcp.approx), 152 FFMA.  Bound: FP32 FMA pipe, not HBM.
// Generic path: any resize factor (incl. identity = masks already upsampled), any C <= 32.
#include <cuda.h>

#include "common.cuh"

namespace mss {

constexpr int M2F_C = 19;          // classes contracted (C of [Q, C+1])
constexpr int M2F_CP = 20;         // padded row of the class-probability table (float4 loads)
constexpr int M2F_MAXQ = 128;
constexpr int QCHUNK = 20;         // queries per TMA stage
constexpr int STAGES = 3;
constexpr int BOX_W = 40, BOX_H = 6;   // low-res patch: cols -4..35 of the tile (34 used): the box must START on a
                                        // 16-byte boundary in global memory (x0 % 4 == 0) or UTMALDG faults
constexpr int BOX_X0 = 4;              // tile column origin = 32*bx - BOX_X0
constexpr int TILE_W = 128, TILE_H = 16;
constexpr int STAGE_FLOATS = QCHUNK * BOX_H * BOX_W;
constexpr int STAGE_BYTES = STAGE_FLOATS * 4;

// ---- class probabilities: softmax over C+1, keep the first C (maskformer_model.py:343 / train_m2f.py:402)
__global__ void m2f_class_probs_kernel(const float *__restrict__ cls, int rows, int C1, int CPAD,
                                       float *__restrict__ probs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *x = cls + (long long)r * C1;
    float m = -INFINITY;
    for (int c = 0; c < C1; c++) m = fmaxf(m, x[c]);
    float s = 0.f;
    for (int c = 0; c < C1; c++) s += expf(x[c] - m);
    float *p = probs + (long long)r * CPAD;
    for (int c = 0; c < CPAD; c++) p[c] = (c < C1 - 1) ? expf(x[c] - m) / s : 0.f;
}

// torch area_pixel_compute_source_index, align_corners=False, non-cubic
__device__ __forceinline__ void src_index_ac0(int dst, float scale, int in_size, int &i0, int &i1, float &l0,
                                              float &l1) {
    float src = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
    i0 = min((int)src, in_size - 1);
    i1 = min(i0 + 1, in_size - 1);
    l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
    l0 = 1.f - l1;
}

__device__ __forceinline__ float sigmoid_fast(float x) {
    // 1 / (1 + 2^(-x log2 e)); ex2.approx + rcp.approx: rel. error ~2^-22, far inside the 1e-5 bar
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}

// ---- mbarrier / TMA primitives ---------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

struct M2FOut {
    float *semseg;                 // [B][...][Hc][Wc], first C planes per image, or null
    long long semseg_bstride;      // elements between images
    float *anomaly;                // [B][Hc][Wc] or null
    float *extra;                  // kept-query planes or null
    long long extra_bstride;
    const int *keep_slot;          // [B][Q]: slot of query q among the kept ones, or -1 (null = none kept)
    const float *keep_score;       // [B][Q]
    int Hc, Wc;
};

// ---- fast path: exact x4, TMA staged ------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
m2f_fused_x4_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ probs, int Q, int h, int w,
                    M2FOut out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_tile = reinterpret_cast<float *>(smem_raw);                               // [STAGES][QCHUNK][6][40]
    float *s_probs = s_tile + STAGES * STAGE_FLOATS;                                   // [Q][20]
    int *s_keep = reinterpret_cast<int *>(s_probs + M2F_MAXQ * M2F_CP);                // [Q]
    float *s_kscore = reinterpret_cast<float *>(s_keep + M2F_MAXQ);                    // [Q]
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_kscore + M2F_MAXQ);              // [STAGES]

    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int b = blockIdx.z;
    const int sx0 = blockIdx.x * (TILE_W / 4) - BOX_X0, sy0 = blockIdx.y * (TILE_H / 4) - 1;
    const int n_chunks = (Q + QCHUNK - 1) / QCHUNK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&s_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int c = 0; c < STAGES && c < n_chunks; c++) {
            mbar_expect_tx(&s_full[c], STAGE_BYTES);
            tma_load_3d(s_tile + c * STAGE_FLOATS, &tmap, &s_full[c], sx0, sy0, b * Q + c * QCHUNK);
        }
    }
    // class probabilities + keep table for this image
    for (int i = tid; i < Q * M2F_CP; i += 256) s_probs[i] = __ldg(probs + (long long)b * Q * M2F_CP + i);
    const bool has_extra = out.keep_slot != nullptr && out.extra != nullptr;
    for (int i = tid; i < Q; i += 256) {
        s_keep[i] = has_extra ? out.keep_slot[(long long)b * Q + i] : -1;
        s_kscore[i] = has_extra ? out.keep_score[(long long)b * Q + i] : 0.f;
    }

    // geometry of this thread's 2 x 4 output block
    const int x4 = blockIdx.x * TILE_W + tx * 4, y2 = blockIdx.y * TILE_H + ty * 2;
    // Taps: output cols 0,1 share the source pair of col 0, cols 2,3 the pair of col 2; both output rows
    // share the source-row pair of row y2.  For an exact x4 resize these pairs are (j-1, j) / (j, j+1)
    // in the interior, and at the borders torch's own clamping (i0 = 0 with l1 == 0 on the left/top,
    // i1 = min(i0+1, in-1) on the right/bottom) yields the same shared pairs.
    int rA, rB, c01a, c01b, c23a, c23b;
    float wx0[4], wx1[4], wy0[2], wy1[2];
    {
        int t0, t1;
        src_index_ac0(x4 + 0, 0.25f, w, c01a, c01b, wx0[0], wx1[0]);
        src_index_ac0(x4 + 1, 0.25f, w, t0, t1, wx0[1], wx1[1]);
        src_index_ac0(x4 + 2, 0.25f, w, c23a, c23b, wx0[2], wx1[2]);
        src_index_ac0(x4 + 3, 0.25f, w, t0, t1, wx0[3], wx1[3]);
        src_index_ac0(y2 + 0, 0.25f, h, rA, rB, wy0[0], wy1[0]);
        src_index_ac0(y2 + 1, 0.25f, h, t0, t1, wy0[1], wy1[1]);
    }
    // smem-local tap offsets (patch origin is (sx0, sy0); all taps are inside the 34 x 6 patch)
    const int o_r0 = (rA - sy0) * BOX_W, o_r1 = (rB - sy0) * BOX_W;
    const int o01a = c01a - sx0, o01b = c01b - sx0, o23a = c23a - sx0, o23b = c23b - sx0;

    float acc[8][M2F_C];
#pragma unroll
    for (int p = 0; p < 8; p++)
#pragma unroll
        for (int c = 0; c < M2F_C; c++) acc[p][c] = 0.f;

    const bool in_crop_row0 = y2 < out.Hc, in_crop_row1 = y2 + 1 < out.Hc;
    __syncthreads();   // s_probs / s_keep visible

    for (int ch = 0; ch < n_chunks; ch++) {
        const int s = ch % STAGES;
        mbar_wait(&s_full[s], (ch / STAGES) & 1);
        const float *tile = s_tile + s * STAGE_FLOATS;
        const int q_lo = ch * QCHUNK, q_n = min(QCHUNK, Q - q_lo);
#pragma unroll 2
        for (int qq = 0; qq < q_n; qq++) {
            const float *t = tile + qq * (BOX_H * BOX_W);
            const float a0 = t[o_r0 + o01a], b0 = t[o_r0 + o01b], c0 = t[o_r0 + o23a], d0 = t[o_r0 + o23b];
            const float a1 = t[o_r1 + o01a], b1 = t[o_r1 + o01b], c1 = t[o_r1 + o23a], d1 = t[o_r1 + o23b];
            float hx0[4], hx1[4];
            hx0[0] = wx0[0] * a0 + wx1[0] * b0; hx0[1] = wx0[1] * a0 + wx1[1] * b0;
            hx0[2] = wx0[2] * c0 + wx1[2] * d0; hx0[3] = wx0[3] * c0 + wx1[3] * d0;
            hx1[0] = wx0[0] * a1 + wx1[0] * b1; hx1[1] = wx0[1] * a1 + wx1[1] * b1;
            hx1[2] = wx0[2] * c1 + wx1[2] * d1; hx1[3] = wx0[3] * c1 + wx1[3] * d1;
            float sg[8];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                sg[j] = sigmoid_fast(wy0[0] * hx0[j] + wy1[0] * hx1[j]);
                sg[4 + j] = sigmoid_fast(wy0[1] * hx0[j] + wy1[1] * hx1[j]);
            }
            const int q = q_lo + qq;
            const float4 *pq = reinterpret_cast<const float4 *>(s_probs + q * M2F_CP);
            float pr[M2F_CP];
#pragma unroll
            for (int v = 0; v < M2F_CP / 4; v++) {
                float4 f = pq[v];
                pr[4 * v] = f.x; pr[4 * v + 1] = f.y; pr[4 * v + 2] = f.z; pr[4 * v + 3] = f.w;
            }
#pragma unroll
            for (int p = 0; p < 8; p++)
#pragma unroll
                for (int c = 0; c < M2F_C; c++) acc[p][c] = fmaf(pr[c], sg[p], acc[p][c]);

            const int slot = s_keep[q];
            if (slot >= 0) {   // warp-uniform, rare: maskformer_model.py:346-352
                const float sc = s_kscore[q];
                float *e = out.extra + (long long)b * out.extra_bstride + (long long)slot * out.Hc * out.Wc;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (x4 + j < out.Wc) {
                        if (in_crop_row0) e[(long long)y2 * out.Wc + x4 + j] = sc * sg[j];
                        if (in_crop_row1) e[(long long)(y2 + 1) * out.Wc + x4 + j] = sc * sg[4 + j];
                    }
                }
            }
        }
        __syncthreads();   // every warp is done with stage s
        if (tid == 0 && ch + STAGES < n_chunks) {
            mbar_expect_tx(&s_full[s], STAGE_BYTES);
            tma_load_3d(s_tile + s * STAGE_FLOATS, &tmap, &s_full[s], sx0, sy0, b * Q + (ch + STAGES) * QCHUNK);
        }
    }

    // epilogue: crop + stores
    const bool vec = (out.Wc % 4 == 0) && (x4 + 3 < out.Wc);
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int y = y2 + r;
        if (y >= out.Hc) continue;
        if (out.semseg) {
            float *base = out.semseg + (long long)b * out.semseg_bstride + (long long)y * out.Wc + x4;
#pragma unroll
            for (int c = 0; c < M2F_C; c++) {
                float *d = base + (long long)c * out.Hc * out.Wc;
                if (vec) {
                    stg_stream_f4(d, make_float4(acc[4 * r][c], acc[4 * r + 1][c], acc[4 * r + 2][c], acc[4 * r + 3][c]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (x4 + j < out.Wc) d[j] = acc[4 * r + j][c];
                }
            }
        }
        if (out.anomaly) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float m = acc[4 * r + j][0];
#pragma unroll
                for (int c = 1; c < M2F_C; c++) m = fmaxf(m, acc[4 * r + j][c]);
                o[j] = 1.0f - m;
            }
            float *d = out.anomaly + ((long long)b * out.Hc + y) * out.Wc + x4;
            if (vec) {
                stg_stream_f4(d, make_float4(o[0], o[1], o[2], o[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (x4 + j < out.Wc) d[j] = o[j];
            }
        }
    }
}

// ---- generic path: any (h,w)->(Hp,Wp), C <= 32; thread = 4 adjacent pixels of one row -------------------
constexpr int GEN_MAXC = 32;

template <bool IDENTITY>
__global__ void __launch_bounds__(128)
m2f_generic_kernel(const float *__restrict__ masks, const float *__restrict__ probs, int Q, int C, int CPAD, int h,
                   int w, int Hp, int Wp, float sh, float sw, M2FOut out) {
    const int b = blockIdx.z;
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x4 >= out.Wc || y >= out.Hc) return;
    int y0, y1, xa[4], xb[4];
    float ly0, ly1, lx0[4], lx1[4];
    src_index_ac0(y, sh, h, y0, y1, ly0, ly1);
#pragma unroll
    for (int j = 0; j < 4; j++) src_index_ac0(min(x4 + j, Wp - 1), sw, w, xa[j], xb[j], lx0[j], lx1[j]);
    float acc[4][GEN_MAXC];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int c = 0; c < GEN_MAXC; c++) acc[j][c] = 0.f;
    const float *mb = masks + (long long)b * Q * h * w;
    const float *pb = probs + (long long)b * Q * CPAD;
    const bool has_extra = out.keep_slot != nullptr && out.extra != nullptr;
    const bool vec_in = IDENTITY && (w % 4 == 0) && (x4 + 3 < w);
    for (int q = 0; q < Q; q++) {
        const float *m = mb + (long long)q * h * w;
        float sg[4];
        if (IDENTITY) {
            // (Hp,Wp) == (h,w): src == dst, l1 == 0, torch's 4-tap formula returns the tap itself
            if (vec_in) {
                float4 v = ldg_stream_f4(m + (long long)y * w + x4);
                sg[0] = v.x; sg[1] = v.y; sg[2] = v.z; sg[3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) sg[j] = __ldg(m + (long long)y * w + min(x4 + j, w - 1));
            }
        } else {
            const float *r0 = m + (long long)y0 * w, *r1 = m + (long long)y1 * w;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float a = __ldg(r0 + xa[j]), bb = __ldg(r0 + xb[j]), c = __ldg(r1 + xa[j]), d = __ldg(r1 + xb[j]);
                sg[j] = ly0 * (lx0[j] * a + lx1[j] * bb) + ly1 * (lx0[j] * c + lx1[j] * d);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) sg[j] = sigmoid_fast(sg[j]);
        const float *pq = pb + (long long)q * CPAD;
#pragma unroll
        for (int c = 0; c < GEN_MAXC; c++) {
            if (c < C) {
                const float p = __ldg(pq + c);
#pragma unroll
                for (int j = 0; j < 4; j++) acc[j][c] = fmaf(p, sg[j], acc[j][c]);
            }
        }
        if (has_extra) {
            const int slot = out.keep_slot[(long long)b * Q + q];
            if (slot >= 0) {
                const float sc = out.keep_score[(long long)b * Q + q];
                float *e = out.extra + (long long)b * out.extra_bstride + (long long)slot * out.Hc * out.Wc +
                           (long long)y * out.Wc + x4;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (x4 + j < out.Wc) e[j] = sc * sg[j];
            }
        }
    }
    if (out.semseg) {
        float *base = out.semseg + (long long)b * out.semseg_bstride + (long long)y * out.Wc + x4;
#pragma unroll
        for (int c = 0; c < GEN_MAXC; c++) {
            if (c < C) {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (x4 + j < out.Wc) base[(long long)c * out.Hc * out.Wc + j] = acc[j][c];
            }
        }
    }
    if (out.anomaly) {
        float *d = out.anomaly + ((long long)b * out.Hc + y) * out.Wc + x4;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float mx = acc[j][0];
#pragma unroll
            for (int c = 1; c < GEN_MAXC; c++)
                if (c < C) mx = fmaxf(mx, acc[j][c]);
            if (x4 + j < out.Wc) d[j] = 1.0f - mx;
        }
    }
}

// keep_idx (compact list per image) -> keep_slot (per query), on device
__global__ void m2f_keep_slots_kernel(const int *__restrict__ keep_idx, const int *__restrict__ keep_count, int Q,
                                      int *__restrict__ keep_slot) {
    const int b = blockIdx.x;
    for (int q = threadIdx.x; q < Q; q += blockDim.x) keep_slot[b * Q + q] = -1;
    __syncthreads();
    const int n = keep_count[b];
    for (int k = threadIdx.x; k < n; k += blockDim.x) keep_slot[b * Q + keep_idx[b * Q + k]] = k;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::atomic<bool> tried{false};
    if (!tried.load()) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        tried.store(true);
    }
    return fn;
}

}  // namespace mss

#include "m2f_mma.cuh"
#include "m2f_tc5.cuh"
#include "m2f_tc5q.cuh"

using namespace mss;

extern "C" size_t mss_m2f_workspace_bytes(int64_t B, int Q, int C) {
    const int CPAD = (C + 3) / 4 * 4;
    return align_up((size_t)B * Q * CPAD * 4, 256) + align_up((size_t)B * Q * 4, 256) +
           2 * align_up((size_t)B * T5_B_FLOATS * 4, 256) + 512;      // T5_B_FLOATS >= MM_QPAD * MM_NPAD
}

extern "C" int mss_m2f_semantic_inference(const float *cls_logits, const float *mask_logits, int64_t B, int Q, int C,
                                             int h, int w, int Hp, int Wp, int Hc, int Wc, float *semseg,
                                             int64_t semseg_batch_stride, float *anomaly, const int32_t *keep_idx,
                                             const float *keep_score, const int32_t *keep_count, float *extra,
                                             int64_t extra_batch_stride, void *workspace, size_t workspace_bytes,
                                             unsigned flags, void *stream) {
    const bool force_generic = (flags & MSS_M2F_FORCE_GENERIC) != 0;
    MSS_REQUIRE(cls_logits && mask_logits && workspace, "mss_m2f_semantic_inference: null pointer");
    MSS_REQUIRE(B >= 0 && Q >= 1 && Q <= M2F_MAXQ && C >= 1 && C <= GEN_MAXC, "mss_m2f_semantic_inference: need 1<=Q<=%d, 1<=C<=%d", M2F_MAXQ, GEN_MAXC);
    MSS_REQUIRE(h > 0 && w > 0 && Hp > 0 && Wp > 0 && Hc > 0 && Wc > 0 && Hc <= Hp && Wc <= Wp,
                "mss_m2f_semantic_inference: bad sizes h=%d w=%d Hp=%d Wp=%d Hc=%d Wc=%d", h, w, Hp, Wp, Hc, Wc);
    MSS_REQUIRE(semseg || anomaly, "mss_m2f_semantic_inference: no output requested");
    MSS_REQUIRE(!semseg || semseg_batch_stride >= (int64_t)C * Hc * Wc, "mss_m2f_semantic_inference: semseg_batch_stride too small");
    if (B == 0) return MSS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int CPAD = (C + 3) / 4 * 4;
    Carver cv(workspace, workspace_bytes);
    float *probs = cv.take<float>((size_t)B * Q * CPAD);
    int *keep_slot = cv.take<int>((size_t)B * Q);
    static_assert(T5_B_FLOATS >= MM_QPAD * MM_NPAD, "split class-probability tables share one workspace slot");
    float *p_hi = cv.take<float>((size_t)B * T5_B_FLOATS);
    float *p_lo = cv.take<float>((size_t)B * T5_B_FLOATS);
    if (!cv.ok()) {
        set_error("mss_m2f_semantic_inference: workspace too small (%zu < %zu)", workspace_bytes, mss_m2f_workspace_bytes(B, Q, C));
        return MSS_ERR_WORKSPACE;
    }
    const int rows = (int)(B * Q);
    // softmax(cls)[:, :C] in plain [row][CPAD] form: only the FFMA / generic kernels read it (the tensor-core paths build
    // their own pre-split tables below), so it is launched where it is needed -- at batch 1 (exps/M2F.yaml:21) every
    // launch in front of the main kernel is a visible fraction of the call
    auto launch_probs = [&]() -> int {
        m2f_class_probs_kernel<<<(rows + 127) / 128, 128, 0, st>>>(cls_logits, rows, C + 1, CPAD, probs);
        MSS_CHECK_LAUNCH();
        return MSS_OK;
    };
    const bool has_extra = keep_idx && keep_score && keep_count && extra;
    if (has_extra) {
        m2f_keep_slots_kernel<<<(unsigned)B, 128, 0, st>>>(keep_idx, keep_count, Q, keep_slot);
        MSS_CHECK_LAUNCH();
    }
    M2FOut out{semseg, semseg_batch_stride, anomaly, has_extra ? extra : nullptr, extra_batch_stride,
               has_extra ? keep_slot : nullptr, keep_score, Hc, Wc};

    const bool x4 = !force_generic && C == M2F_C && Hp == 4 * h && Wp == 4 * w && (w % 4 == 0) && h >= 2 && w >= 2 &&
                    ((uintptr_t)mask_logits % 16 == 0) && B <= 65535;
    EncodeTiledFn enc = x4 ? get_encode_fn() : nullptr;
    const bool use_mma = x4 && enc && !(flags & MSS_M2F_FORCE_FFMA) && Q <= MM_QPAD - 4;
    const bool use_tc5 = use_mma && !(flags & MSS_M2F_FORCE_MMASYNC);
    const bool use_tc5q = use_tc5 && !(flags & MSS_M2F_FORCE_TC5_PIXEL) && (Q % 4 == 0);
    if (x4 && enc) {
        CUtensorMap tmap;
        cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)(B * Q)};
        cuuint64_t gstr[2] = {(cuuint64_t)w * 4, (cuuint64_t)w * h * 4};
        cuuint32_t box[3] = {BOX_W, BOX_H, QCHUNK};
        if (use_mma) { box[0] = MM_BOX_W; box[1] = MM_BOX_H; box[2] = MM_QPAD; }
        if (use_tc5) { box[0] = T5_BOX_W; box[1] = T5_BOX_H; box[2] = T5_K; }
        if (use_tc5q) { box[0] = TQ_BOX_W; box[1] = TQ_BOX_H; box[2] = T5_K; }
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)mask_logits, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
            return MSS_ERR_CUDA;
        }
        if (use_tc5) {
            m2f_class_probs_umma_kernel<<<(unsigned)((B * T5_K + 127) / 128), 128, 0, st>>>(cls_logits, (int)B, Q, C + 1, p_hi, p_lo);
            MSS_CHECK_LAUNCH();
            if (use_tc5q) {
                // debug switch: MSS_M2F_DUP_B=1 keeps two copies of each class-table core matrix instead of LBO = 0
                static const bool dup_b = [] { const char *e = getenv("MSS_M2F_DUP_B"); return e && e[0] == '1'; }();
                static std::atomic<unsigned long long> tq_attr_set{0};
                if (first_use_on_device(tq_attr_set)) {
                    MSS_CHECK_CUDA(cudaFuncSetAttribute(m2f_tc5q_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tq_smem(false)));
                    MSS_CHECK_CUDA(cudaFuncSetAttribute(m2f_tc5q_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tq_smem(false)));
                    MSS_CHECK_CUDA(cudaFuncSetAttribute(m2f_tc5q_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tq_smem(true)));
                    MSS_CHECK_CUDA(cudaFuncSetAttribute(m2f_tc5q_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tq_smem(true)));
                }
                // blocks of 4 pixels start at x = 2 (mod 4): CTA bx covers x in [64 bx - 2, 64 bx + 62)
                dim3 grid((Wc + 2 + TQ_W - 1) / TQ_W, (Hc + TQ_H - 1) / TQ_H, (unsigned)B);
                MSS_REQUIRE(grid.y <= 65535, "mss_m2f_semantic_inference: grid too large");
                if (dup_b) {
                    if (has_extra) m2f_tc5q_kernel<true, true><<<grid, TQ_THREADS, tq_smem(true), st>>>(tmap, p_hi, p_lo, Q, h, w, out);
                    else m2f_tc5q_kernel<false, true><<<grid, TQ_THREADS, tq_smem(true), st>>>(tmap, p_hi, p_lo, Q, h, w, out);
                } else {
                    if (has_extra) m2f_tc5q_kernel<true, false><<<grid, TQ_THREADS, tq_smem(false), st>>>(tmap, p_hi, p_lo, Q, h, w, out);
                    else m2f_tc5q_kernel<false, false><<<grid, TQ_THREADS, tq_smem(false), st>>>(tmap, p_hi, p_lo, Q, h, w, out);
                }
                MSS_CHECK_LAUNCH();
                return MSS_OK;
            }
            static std::atomic<unsigned long long> tc5_attr_set{0};
            if (first_use_on_device(tc5_attr_set)) {
                MSS_CHECK_CUDA(cudaFuncSetAttribute(m2f_tc5_x4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T5_SMEM));
                MSS_CHECK_CUDA(cudaFuncSetAttribute(m2f_tc5_x4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T5_SMEM));
            }
            dim3 grid((Wc + T5_TILE_W - 1) / T5_TILE_W, (Hc + T5_BLOCK_H - 1) / T5_BLOCK_H, (unsigned)B);
            MSS_REQUIRE(grid.y <= 65535, "mss_m2f_semantic_inference: grid too large");
            if (has_extra)
                m2f_tc5_x4_kernel<true><<<grid, T5_THREADS, T5_SMEM, st>>>(tmap, p_hi, p_lo, Q, h, w, out);
            else
                m2f_tc5_x4_kernel<false><<<grid, T5_THREADS, T5_SMEM, st>>>(tmap, p_hi, p_lo, Q, h, w, out);
            MSS_CHECK_LAUNCH();
            return MSS_OK;
        }
        if (use_mma) {
            m2f_class_probs_split_kernel<<<(unsigned)((B * MM_QPAD + 127) / 128), 128, 0, st>>>(cls_logits, (int)B, Q, C + 1,
                                                                                              p_hi, p_lo);
            MSS_CHECK_LAUNCH();
            const size_t smem = (size_t)MM_PATCH_BYTES + 2 * (size_t)MM_QPAD * MM_NPAD * 4 + MM_QPAD * 8 + 8 + 128;
            static std::atomic<unsigned long long> mma_attr_set{0};
            if (first_use_on_device(mma_attr_set)) {
                MSS_CHECK_CUDA(cudaFuncSetAttribute(m2f_mma_x4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            }
            dim3 grid((Wc + MM_TILE_W - 1) / MM_TILE_W, (Hc + MM_TILE_H - 1) / MM_TILE_H, (unsigned)B);
            MSS_REQUIRE(grid.y <= 65535, "mss_m2f_semantic_inference: grid too large");
            m2f_mma_x4_kernel<<<grid, 128, smem, st>>>(tmap, p_hi, p_lo, Q, h, w, out);
            MSS_CHECK_LAUNCH();
            return MSS_OK;
        }
        if (int rc = launch_probs()) return rc;
        const size_t smem = (size_t)STAGES * STAGE_BYTES + (size_t)M2F_MAXQ * M2F_CP * 4 + M2F_MAXQ * 8 + STAGES * 8 + 128;
        static std::atomic<unsigned long long> attr_set{0};
        if (first_use_on_device(attr_set)) {
            MSS_CHECK_CUDA(cudaFuncSetAttribute(m2f_fused_x4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        dim3 grid((Wc + TILE_W - 1) / TILE_W, (Hc + TILE_H - 1) / TILE_H, (unsigned)B);
        m2f_fused_x4_kernel<<<grid, 256, smem, st>>>(tmap, probs, Q, h, w, out);
        MSS_CHECK_LAUNCH();
        return MSS_OK;
    }
    if (int rc = launch_probs()) return rc;
    const bool identity = (Hp == h && Wp == w);
    dim3 grid(((Wc + 3) / 4 + 127) / 128, Hc, (unsigned)B);
    MSS_REQUIRE(Hc <= 65535 && B <= 65535, "mss_m2f_semantic_inference: grid too large");
    const float sh = (float)h / (float)Hp, sw = (float)w / (float)Wp;
    if (identity)
        m2f_generic_kernel<true><<<grid, 128, 0, st>>>(mask_logits, probs, Q, C, CPAD, h, w, Hp, Wp, sh, sw, out);
    else
        m2f_generic_kernel<false><<<grid, 128, 0, st>>>(mask_logits, probs, Q, C, CPAD, h, w, Hp, Wp, sh, sw, out);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}
