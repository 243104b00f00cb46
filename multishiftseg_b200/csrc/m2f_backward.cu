// (Mask2Former half) backward of the fused anomaly score, for the trainer's use of the same path with
// autograd enabled: TrainM2FOOD.get_anomaly_score (train_m2f.py:387-407) is called inside the training step
// (train_m2f.py:443) and its result feeds criterion.loss_ood (modeling/criterion.py:128-187) and the contrastive loss
// (lib/loss.py:119-147).  Forward, per image:
//     P = softmax(cls)[:, :C]                                    [Q, C]
//     U = bilinear(mask_logits -> Hp x Wp, align_corners=False)  [Q, Hp, Wp]      (maskformer_model.py:271-277)
//     S = sigmoid(U);  sem[c, px] = sum_q P[q, c] S[q, px];  score[px] = 1 - max_c sem[c, px]   (cropped to Hc x Wc)
// Backward for g = dL/dscore, with c*(px) = argmax_c sem[c, px] (torch.max routes the gradient to the maximal class):
//     dL/dU[q, px]      = -g[px] P[q, c*(px)] S (1 - S)
//     dL/dmask[q, y, x] = sum over the output pixels whose forward taps include (y, x) of  weight * dL/dU[q, px]
//     dL/dP[q, c]       = -sum_px [c*(px) == c] g[px] S[q, px]
//     dL/dcls[q, j]     = p_j (dP_j - sum_c p_c dP_c),  dP_C := 0 for the dropped "no object" column
// Nothing of size [Q, Hp, Wp] is materialised: S is recomputed from the decoder-resolution masks where it is needed.
//   1. m2f_bwd_argmax_kernel      one thread per pixel: sem over all queries (FFMA), c* as one byte per pixel
//   2. m2f_bwd_masks_kernel       GATHER form of the upsample adjoint (no atomics, deterministic): one thread per
//                                 (decoder-resolution cell, chunk of 8 queries) walks the output pixels of its footprint;
//                                 the tap weights are recomputed with the forward's own index arithmetic, so the adjoint
//                                 is exact for any resize factor, borders and clamps included
//   3. m2f_bwd_cls_partial_kernel per CTA tile of pixels: class-binned sums of -g S in a fixed order (private
//                                 accumulators, warp tree, warps in order), one [Q, C] partial per CTA
//   4. m2f_bwd_cls_final_kernel   partials summed in CTA order, softmax backward; deterministic as well
// These are correctness-first kernels (no tensor cores, no TMA): about 2 ms per 1024 x 2048 image on B200.
#include "common.cuh"

namespace mss {

constexpr int MB_MAXQ = 128, MB_MAXC = 32;

// torch area_pixel_compute_source_index, align_corners=False (same arithmetic as m2f_semantic.cu: src_index_ac0)
__device__ __forceinline__ void mb_src_index(int dst, float scale, int in_size, int &i0, int &i1, float &l0, float &l1) {
    float src = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
    i0 = min((int)src, in_size - 1);
    i1 = min(i0 + 1, in_size - 1);
    l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
    l0 = 1.f - l1;
}
// conservative range of output indices whose taps can touch input index i: src in (i - 1, i + 1), widened by one
__device__ __forceinline__ void mb_out_range(int i, float scale, int out_size, int &lo, int &hi) {
    const float inv = 1.0f / scale;
    lo = max(0, (int)floorf(((float)i - 0.5f) * inv - 0.5f) - 1);
    hi = min(out_size - 1, (int)ceilf(((float)i + 1.5f) * inv - 0.5f) + 1);
}
__device__ __forceinline__ float mb_sigmoid(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}

// softmax(cls)[:, :C] -> probs [rows][MB_MAXC] (zero padded), and the full softmax row for the final step
__global__ void m2f_bwd_probs_kernel(const float *__restrict__ cls, int rows, int C1, float *__restrict__ probs,
                                     float *__restrict__ full) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *x = cls + (long long)r * C1;
    float m = -INFINITY;
    for (int c = 0; c < C1; c++) m = fmaxf(m, x[c]);
    float s = 0.f;
    for (int c = 0; c < C1; c++) s += expf(x[c] - m);
    for (int c = 0; c < MB_MAXC; c++) probs[(long long)r * MB_MAXC + c] = (c < C1 - 1) ? expf(x[c] - m) / s : 0.f;
    for (int c = 0; c < C1; c++) full[(long long)r * C1 + c] = expf(x[c] - m) / s;
}

// ---- 1. argmax class per pixel ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
m2f_bwd_argmax_kernel(const float *__restrict__ masks, const float *__restrict__ probs, int Q, int C, int h, int w, int Hp,
                      int Wp, int Hc, int Wc, float sh, float sw, uint8_t *__restrict__ cstar) {
    extern __shared__ float s_p[];                       // [Q][MB_MAXC]
    const int b = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = threadIdx.x; i < Q * MB_MAXC; i += blockDim.x) s_p[i] = probs[(long long)b * Q * MB_MAXC + i];
    __syncthreads();
    if (x >= Wc) return;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    mb_src_index(y, sh, h, y0, y1, ly0, ly1);
    mb_src_index(x, sw, w, x0, x1, lx0, lx1);
    float acc[MB_MAXC];
#pragma unroll
    for (int c = 0; c < MB_MAXC; c++) acc[c] = 0.f;
    const float *mb = masks + (long long)b * Q * h * w;
    for (int q = 0; q < Q; q++) {
        const float *r0 = mb + ((long long)q * h + y0) * w, *r1 = mb + ((long long)q * h + y1) * w;
        const float u = ly0 * (lx0 * __ldg(r0 + x0) + lx1 * __ldg(r0 + x1)) + ly1 * (lx0 * __ldg(r1 + x0) + lx1 * __ldg(r1 + x1));
        const float s = mb_sigmoid(u);
        const float *pq = s_p + q * MB_MAXC;
#pragma unroll
        for (int c = 0; c < MB_MAXC; c++)
            if (c < C) acc[c] = fmaf(pq[c], s, acc[c]);
    }
    int best = 0;
    float mx = acc[0];
#pragma unroll
    for (int c = 1; c < MB_MAXC; c++)
        if (c < C && acc[c] > mx) { mx = acc[c]; best = c; }           // first maximal index, as torch.max
    cstar[((long long)b * Hc + y) * Wc + x] = (uint8_t)best;
}

// ---- 2. gradient of the decoder-resolution masks (gather form) ---------------------------------------------------
constexpr int MBQ = 8;                                    // queries per thread
__global__ void __launch_bounds__(128)
m2f_bwd_masks_kernel(const float *__restrict__ masks, const float *__restrict__ probs, const float *__restrict__ g,
                     const uint8_t *__restrict__ cstar, int Q, int h, int w, int Hp, int Wp, int Hc, int Wc, float sh, float sw,
                     float *__restrict__ grad_masks) {
    const int b = blockIdx.z;
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    const int q0 = blockIdx.y * MBQ;
    if (cell >= h * w) return;
    const int cy = cell / w, cx = cell - cy * w;
    int Y0, Y1, X0, X1;
    mb_out_range(cy, sh, Hp, Y0, Y1);
    mb_out_range(cx, sw, Wp, X0, X1);
    Y1 = min(Y1, Hc - 1);                                 // pixels outside the crop carry no gradient
    X1 = min(X1, Wc - 1);
    float acc[MBQ];
#pragma unroll
    for (int j = 0; j < MBQ; j++) acc[j] = 0.f;
    const float *mb = masks + (long long)b * Q * h * w;
    const float *pb = probs + (long long)b * Q * MB_MAXC;
    const float *gb = g + (long long)b * Hc * Wc;
    const uint8_t *cb = cstar + (long long)b * Hc * Wc;
    for (int Y = Y0; Y <= Y1; Y++) {
        int y0, y1;
        float ly0, ly1;
        mb_src_index(Y, sh, h, y0, y1, ly0, ly1);
        const float wy = (y0 == cy ? ly0 : 0.f) + (y1 == cy ? ly1 : 0.f);
        if (wy == 0.f) continue;
        for (int X = X0; X <= X1; X++) {
            int x0, x1;
            float lx0, lx1;
            mb_src_index(X, sw, w, x0, x1, lx0, lx1);
            const float wx = (x0 == cx ? lx0 : 0.f) + (x1 == cx ? lx1 : 0.f);
            if (wx == 0.f) continue;
            const float coef = -__ldg(gb + (long long)Y * Wc + X) * wy * wx;
            if (coef == 0.f) continue;
            const int cs = cb[(long long)Y * Wc + X];
#pragma unroll
            for (int j = 0; j < MBQ; j++) {
                const int q = q0 + j;
                if (q < Q) {
                    const float *r0 = mb + ((long long)q * h + y0) * w, *r1 = mb + ((long long)q * h + y1) * w;
                    const float u = ly0 * (lx0 * __ldg(r0 + x0) + lx1 * __ldg(r0 + x1)) +
                                    ly1 * (lx0 * __ldg(r1 + x0) + lx1 * __ldg(r1 + x1));
                    const float s = mb_sigmoid(u);
                    acc[j] = fmaf(coef * __ldg(pb + q * MB_MAXC + cs), s * (1.0f - s), acc[j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < MBQ; j++)
        if (q0 + j < Q) grad_masks[((long long)b * Q + q0 + j) * h * w + cell] = acc[j];
}

// ---- 3. class-binned sums of -g S: one [Q][C] partial per CTA tile of pixels ---------------------------------------
constexpr int MC_PX = 4;                                  // pixels per thread
constexpr int MCT = 256;
__global__ void __launch_bounds__(MCT)
m2f_bwd_cls_partial_kernel(const float *__restrict__ masks, const float *__restrict__ g, const uint8_t *__restrict__ cstar,
                           int Q, int C, int h, int w, int Hc, int Wc, float sh, float sw, float *__restrict__ partial) {
    __shared__ float s_warp[MCT / 32][MB_MAXC];
    const int b = blockIdx.y;
    const long long npx = (long long)Hc * Wc;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // this thread's pixels: geometry, coefficient, class
    int y0[MC_PX], y1[MC_PX], x0[MC_PX], x1[MC_PX], cs[MC_PX];
    float ly1[MC_PX], lx1[MC_PX], coef[MC_PX];
#pragma unroll
    for (int j = 0; j < MC_PX; j++) {
        const long long p = ((long long)blockIdx.x * MCT + threadIdx.x) * MC_PX + j;
        const bool in = p < npx;
        const int Y = in ? (int)(p / Wc) : 0, X = in ? (int)(p - (long long)Y * Wc) : 0;
        float l0;
        mb_src_index(Y, sh, h, y0[j], y1[j], l0, ly1[j]);
        mb_src_index(X, sw, w, x0[j], x1[j], l0, lx1[j]);
        coef[j] = in ? -__ldg(g + (long long)b * npx + p) : 0.f;
        cs[j] = in ? (int)cstar[(long long)b * npx + p] : 0;
    }
    const float *mb = masks + (long long)b * Q * h * w;
    float *out = partial + ((long long)b * gridDim.x + blockIdx.x) * Q * MB_MAXC;
    for (int q = 0; q < Q; q++) {
        float acc[MB_MAXC];
#pragma unroll
        for (int c = 0; c < MB_MAXC; c++) acc[c] = 0.f;
#pragma unroll
        for (int j = 0; j < MC_PX; j++) {
            const float *r0 = mb + ((long long)q * h + y0[j]) * w, *r1 = mb + ((long long)q * h + y1[j]) * w;
            const float ly0 = 1.f - ly1[j], lx0 = 1.f - lx1[j];
            const float u = ly0 * (lx0 * __ldg(r0 + x0[j]) + lx1[j] * __ldg(r0 + x1[j])) +
                            ly1[j] * (lx0 * __ldg(r1 + x0[j]) + lx1[j] * __ldg(r1 + x1[j]));
            const float v = coef[j] * mb_sigmoid(u);
#pragma unroll
            for (int c = 0; c < MB_MAXC; c++)
                if (c < C) acc[c] += (cs[j] == c) ? v : 0.f;
        }
        // fixed-order reduction: xor tree inside the warp, then the warps in order
#pragma unroll
        for (int c = 0; c < MB_MAXC; c++) {
            if (c < C) {
                float t = acc[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0) s_warp[warp][c] = t;
            }
        }
        __syncthreads();
        if (threadIdx.x < MB_MAXC) {
            float t = 0.f;
            if ((int)threadIdx.x < C)
                for (int wv = 0; wv < MCT / 32; wv++) t += s_warp[wv][threadIdx.x];
            out[q * MB_MAXC + threadIdx.x] = t;
        }
        __syncthreads();
    }
}

// ---- 4. partials -> dP -> softmax backward ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
m2f_bwd_cls_final_kernel(const float *__restrict__ partial, int n_part, const float *__restrict__ full, int Q, int C,
                         float *__restrict__ grad_cls) {
    const int b = blockIdx.y, q = blockIdx.x, lane = threadIdx.x;
    float dP = 0.f;
    if (lane < C)
        for (int t = 0; t < n_part; t++) dP += partial[(((long long)b * n_part + t) * Q + q) * MB_MAXC + lane];
    const float *p = full + ((long long)b * Q + q) * (C + 1);
    const float pj = (lane <= C) ? p[lane] : 0.f;
    float dot = (lane < C) ? pj * dP : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane <= C) grad_cls[((long long)b * Q + q) * (C + 1) + lane] = pj * ((lane < C ? dP : 0.f) - dot);
}

static float mb_scale(int in, int out) { return (float)in / (float)out; }      // torch: align_corners=False, no scale_factor

}  // namespace mss

using namespace mss;

static long long cls_tiles(long long npx) { return (npx + (long long)MCT * MC_PX - 1) / ((long long)MCT * MC_PX); }

extern "C" size_t mss_m2f_anomaly_backward_workspace_bytes(int64_t B, int Q, int C, int Hc, int Wc) {
    if (B < 0) B = 0;
    const long long npx = (long long)Hc * Wc;
    return align_up((size_t)B * Q * MB_MAXC * 4, 256) + align_up((size_t)B * Q * (C + 1) * 4, 256) + align_up((size_t)B * npx, 256) +
           align_up((size_t)B * cls_tiles(npx) * Q * MB_MAXC * 4, 256) + 1024;
}

extern "C" int mss_m2f_anomaly_backward(const float *cls_logits, const float *mask_logits, const float *grad_score, int64_t B,
                                        int Q, int C, int h, int w, int Hp, int Wp, int Hc, int Wc, float *grad_cls,
                                        float *grad_masks, void *workspace, size_t workspace_bytes, void *stream) {
    MSS_REQUIRE(cls_logits && mask_logits && grad_score && grad_cls && grad_masks && workspace,
                "mss_m2f_anomaly_backward: null pointer");
    MSS_REQUIRE(B >= 0 && Q >= 1 && Q <= MB_MAXQ && C >= 1 && C < MB_MAXC, "mss_m2f_anomaly_backward: need 1<=Q<=%d, 1<=C<%d",
                MB_MAXQ, MB_MAXC);
    MSS_REQUIRE(h > 0 && w > 0 && Hp > 0 && Wp > 0 && Hc > 0 && Wc > 0 && Hc <= Hp && Wc <= Wp,
                "mss_m2f_anomaly_backward: bad sizes h=%d w=%d Hp=%d Wp=%d Hc=%d Wc=%d", h, w, Hp, Wp, Hc, Wc);
    if (B == 0) return MSS_OK;
    MSS_REQUIRE(B <= 65535 && Hc <= 65535, "mss_m2f_anomaly_backward: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    const long long npx = (long long)Hc * Wc, tiles = cls_tiles(npx);
    Carver cv(workspace, workspace_bytes);
    float *probs = cv.take<float>((size_t)B * Q * MB_MAXC);
    float *full = cv.take<float>((size_t)B * Q * (C + 1));
    uint8_t *cstar = cv.take<uint8_t>((size_t)B * npx);
    float *partial = cv.take<float>((size_t)B * tiles * Q * MB_MAXC);
    if (!cv.ok()) {
        set_error("mss_m2f_anomaly_backward: workspace too small (%zu < %zu)", workspace_bytes,
                  mss_m2f_anomaly_backward_workspace_bytes(B, Q, C, Hc, Wc));
        return MSS_ERR_WORKSPACE;
    }
    const float sh = mb_scale(h, Hp), sw = mb_scale(w, Wp);
    const int rows = (int)(B * Q);
    m2f_bwd_probs_kernel<<<(rows + 127) / 128, 128, 0, st>>>(cls_logits, rows, C + 1, probs, full);
    MSS_CHECK_LAUNCH();
    m2f_bwd_argmax_kernel<<<dim3((Wc + 127) / 128, Hc, (unsigned)B), 128, (size_t)Q * MB_MAXC * 4, st>>>(
        mask_logits, probs, Q, C, h, w, Hp, Wp, Hc, Wc, sh, sw, cstar);
    MSS_CHECK_LAUNCH();
    m2f_bwd_masks_kernel<<<dim3((h * w + 127) / 128, (Q + MBQ - 1) / MBQ, (unsigned)B), 128, 0, st>>>(
        mask_logits, probs, grad_score, cstar, Q, h, w, Hp, Wp, Hc, Wc, sh, sw, grad_masks);
    MSS_CHECK_LAUNCH();
    m2f_bwd_cls_partial_kernel<<<dim3((unsigned)tiles, (unsigned)B), MCT, 0, st>>>(mask_logits, grad_score, cstar, Q, C, h, w, Hc, Wc,
                                                                                   sh, sw, partial);
    MSS_CHECK_LAUNCH();
    m2f_bwd_cls_final_kernel<<<dim3(Q, (unsigned)B), 32, 0, st>>>(partial, (int)tiles, full, Q, C, grad_cls);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}
