// (DeepLab half) training-time reuse of the scoring path: the reference's trainer runs the SAME
// energy_func + Upsample with autograd enabled (train_deeplab.py:197-198: `anomaly_score, logit = self.model(img)`,
// `loss = self.criterion(logit, anomaly_score, target)`; lib/loss.py:34-147 consumes the score map), so a drop-in for
// deepv3.py:251-253 / :283 that is used for training needs the backward of both ops:
//     s = -(logsumexp_c x)                      ds/dx_c = -softmax(x)_c                         (energy_func)
//     Y = bilinear(X, align_corners)            dL/dX = A^T dL/dY, A = the forward's tap matrix  (mynn.Upsample)
// Kernels (all HBM-bound elementwise / gather work, fp32):
//   * energy_backward_vec4_kernel<19>: thread = 4 adjacent pixels; 19 x 128-bit loads, softmax in registers,
//     19 x 128-bit streaming stores of -p_c * g: 76 + 4 B read, 76 B written per pixel;
//   * upsample_bilinear_backward_kernel: GATHER form of the adjoint (no atomics, deterministic): an input pixel (y, x)
//     collects g(Y, X) * wy * wx from every output pixel whose forward taps include it; the taps are recomputed with the
//     forward's own index arithmetic (upsample.cu: src_index), so A^T is exact for either align_corners mode, borders
//     and clamps included;
//   * deeplab_anomaly_backward_kernel: both fused for deepv3.py:283 -- one thread per head-resolution pixel gathers the
//     upsampled gradient, then writes -softmax(dec2) * g for the C channels (lane = pixel: coalesced per plane).
#include "common.cuh"

namespace mss {

// same arithmetic as upsample.cu (kept identical on purpose: the backward must be the adjoint of THAT forward)
__device__ __forceinline__ void bw_src_index(int dst, float scale, int align, int in_size, int &i0, int &i1, float &l0,
                                             float &l1) {
    float src = align ? scale * (float)dst : fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
    i0 = min((int)src, in_size - 1);
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
    l0 = 1.f - l1;
}

// conservative range of output indices whose taps can touch input index i: src in (i - 1, i + 1)
__device__ __forceinline__ void bw_out_range(int i, float scale, int align, int out_size, int &lo, int &hi) {
    if (scale <= 0.f) { lo = 0; hi = out_size - 1; return; }            // out_size == 1 with align_corners
    const float inv = 1.0f / scale;
    const float off = align ? 0.f : 0.5f;
    // src = scale * (dst + off) - off  =>  dst = (src + off) / scale - off ; widen by one pixel on both sides
    lo = max(0, (int)floorf(((float)i - 1.f + off) * inv - off) - 1);
    hi = min(out_size - 1, (int)ceilf(((float)i + 1.f + off) * inv - off) + 1);
}

// weight with which output index `dst` reads input index `i` in the forward (0 if it does not)
__device__ __forceinline__ float bw_tap_weight(int dst, int i, float scale, int align, int in_size) {
    int a0, a1;
    float l0, l1;
    bw_src_index(dst, scale, align, in_size, a0, a1, l0, l1);
    return (a0 == i ? l0 : 0.f) + (a1 == i ? l1 : 0.f);
}

// sum over the output pixels (Y, X) of plane `g` whose forward taps include input pixel (y, x).  The column weights
// are computed once per thread (registers, ranges of up to BW_R candidates -- any upsampling factor up to ~2.5) instead
// of once per candidate row; the order of the additions is the same in both forms.
constexpr int BW_R = 8;
__device__ __forceinline__ float bw_gather(const float *__restrict__ g, int y, int x, int h, int w, int H, int W, float sh,
                                           float sw, int align) {
    int Y0, Y1, X0, X1;
    bw_out_range(y, sh, align, H, Y0, Y1);
    bw_out_range(x, sw, align, W, X0, X1);
    float acc = 0.f;
    if (X1 - X0 < BW_R) {
        float wxv[BW_R];
#pragma unroll
        for (int k = 0; k < BW_R; k++) wxv[k] = (X0 + k <= X1) ? bw_tap_weight(X0 + k, x, sw, align, w) : 0.f;
        for (int Y = Y0; Y <= Y1; Y++) {
            const float wy = bw_tap_weight(Y, y, sh, align, h);
            if (wy == 0.f) continue;
            const float *row = g + (long long)Y * W + X0;
            float racc = 0.f;
#pragma unroll
            for (int k = 0; k < BW_R; k++)
                if (wxv[k] != 0.f) racc += wxv[k] * __ldg(row + k);
            acc += wy * racc;
        }
        return acc;
    }
    for (int Y = Y0; Y <= Y1; Y++) {
        const float wy = bw_tap_weight(Y, y, sh, align, h);
        if (wy == 0.f) continue;
        const float *row = g + (long long)Y * W;
        float racc = 0.f;
        for (int X = X0; X <= X1; X++) {
            const float wx = bw_tap_weight(X, x, sw, align, w);
            if (wx != 0.f) racc += wx * __ldg(row + X);
        }
        acc += wy * racc;
    }
    return acc;
}

__global__ void __launch_bounds__(256)
upsample_bilinear_backward_kernel(const float *__restrict__ grad_out, int h, int w, float *__restrict__ grad_in, int H,
                                  int W, float sh, float sw, int align, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over NC * h * w
    if (i >= n) return;
    const int x = (int)(i % w);
    const long long r = i / w;
    const int y = (int)(r % h);
    const long long nc = r / h;
    grad_in[i] = bw_gather(grad_out + nc * (long long)H * W, y, x, h, w, H, W, sh, sw, align);
}

// grad_logits[b, c, p] = -softmax(logits[b, :, p])_c * grad_score[b, p]
template <int C>
__global__ void __launch_bounds__(256)
energy_backward_vec4_kernel(const float *__restrict__ logits, const float *__restrict__ grad_score,
                            float *__restrict__ grad_logits, long long HW, long long n_quads, long long quads_per_image) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_quads) return;
    const long long b = q / quads_per_image, p = (q - b * quads_per_image) * 4;
    const float *src = logits + b * C * HW + p;
    float4 v[C];
#pragma unroll
    for (int c = 0; c < C; c++) v[c] = ldg_stream_f4(src + (long long)c * HW);
    const float4 g = ldg_stream_f4(grad_score + b * HW + p);
    float4 m = v[0];
#pragma unroll
    for (int c = 1; c < C; c++) {
        m.x = fmaxf(m.x, v[c].x); m.y = fmaxf(m.y, v[c].y); m.z = fmaxf(m.z, v[c].z); m.w = fmaxf(m.w, v[c].w);
    }
    // torch.logsumexp's guard: an infinite maximum is replaced by 0
    m.x = isinf(m.x) ? 0.f : m.x; m.y = isinf(m.y) ? 0.f : m.y; m.z = isinf(m.z) ? 0.f : m.z; m.w = isinf(m.w) ? 0.f : m.w;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < C; c++) {
        v[c].x = expf(v[c].x - m.x); v[c].y = expf(v[c].y - m.y); v[c].z = expf(v[c].z - m.z); v[c].w = expf(v[c].w - m.w);
        s.x += v[c].x; s.y += v[c].y; s.z += v[c].z; s.w += v[c].w;
    }
    const float4 k = make_float4(-g.x / s.x, -g.y / s.y, -g.z / s.z, -g.w / s.w);
    float *dst = grad_logits + b * C * HW + p;
#pragma unroll
    for (int c = 0; c < C; c++)
        stg_stream_f4(dst + (long long)c * HW, make_float4(v[c].x * k.x, v[c].y * k.y, v[c].z * k.z, v[c].w * k.w));
}

// any C, any alignment: one thread per pixel, channels strided (lane = pixel: coalesced per plane).  With `up` the
// per-pixel gradient is first gathered from a [B, H, W] map through the adjoint of the bilinear upsample.
__global__ void __launch_bounds__(256)
energy_backward_generic_kernel(const float *__restrict__ logits, const float *__restrict__ grad, float *__restrict__ grad_logits,
                               int C, int h, int w, long long n, int up, int H, int W, float sh, float sw, int align) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over B * h * w
    if (i >= n) return;
    const long long hw = (long long)h * w;
    const long long b = i / hw, p = i - b * hw;
    float g;
    if (up) g = bw_gather(grad + b * (long long)H * W, (int)(p / w), (int)(p % w), h, w, H, W, sh, sw, align);
    else g = grad[i];
    const float *src = logits + b * C * hw + p;
    float m = -INFINITY;
    for (int c = 0; c < C; c++) m = fmaxf(m, __ldg(src + (long long)c * hw));
    const float ms = isinf(m) ? 0.f : m;                       // torch.logsumexp's guard
    float s = 0.f;
    for (int c = 0; c < C; c++) s += expf(__ldg(src + (long long)c * hw) - ms);
    const float k = -g / s;
    float *dst = grad_logits + b * C * hw + p;
    for (int c = 0; c < C; c++) dst[(long long)c * hw] = expf(__ldg(src + (long long)c * hw) - ms) * k;
}

static float bw_resize_scale(int in, int out, int align) {
    if (align) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    return (float)in / (float)out;
}

}  // namespace mss

using namespace mss;

extern "C" int mss_upsample_bilinear_backward(const float *grad_out, int64_t NC, int h, int w, float *grad_in, int H, int W,
                                              int align_corners, void *stream) {
    MSS_REQUIRE(grad_out && grad_in, "mss_upsample_bilinear_backward: null pointer");
    MSS_REQUIRE(NC >= 0 && h > 0 && w > 0 && H > 0 && W > 0, "mss_upsample_bilinear_backward: bad shape");
    if (NC == 0) return MSS_OK;
    const long long n = (long long)NC * h * w;
    const long long blocks = (n + 255) / 256;
    MSS_REQUIRE(blocks < (1ll << 31), "mss_upsample_bilinear_backward: tensor too large");
    upsample_bilinear_backward_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        grad_out, h, w, grad_in, H, W, bw_resize_scale(h, H, align_corners), bw_resize_scale(w, W, align_corners),
        align_corners ? 1 : 0, n);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

extern "C" int mss_deeplab_energy_backward(const float *logits, const float *grad_score, int64_t B, int C, int64_t HW,
                                           float *grad_logits, void *stream) {
    MSS_REQUIRE(B >= 0 && C >= 1 && HW >= 0, "mss_deeplab_energy_backward: bad shape");
    if (B == 0 || HW == 0) return MSS_OK;
    MSS_REQUIRE(logits && grad_score && grad_logits, "mss_deeplab_energy_backward: null pointer");
    MSS_REQUIRE(HW < (1ll << 31), "mss_deeplab_energy_backward: image too large");
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = C == 19 && HW % 4 == 0 && (((uintptr_t)logits | (uintptr_t)grad_score | (uintptr_t)grad_logits) & 15) == 0;
    if (vec) {
        const long long qpi = HW / 4, n_quads = B * qpi, blocks = (n_quads + 255) / 256;
        MSS_REQUIRE(blocks < (1ll << 31), "mss_deeplab_energy_backward: tensor too large");
        energy_backward_vec4_kernel<19><<<(unsigned)blocks, 256, 0, st>>>(logits, grad_score, grad_logits, HW, n_quads, qpi);
    } else {
        const long long n = B * HW, blocks = (n + 255) / 256;
        MSS_REQUIRE(blocks < (1ll << 31), "mss_deeplab_energy_backward: tensor too large");
        energy_backward_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(logits, grad_score, grad_logits, C, 1, (int)HW, n, 0,
                                                                        1, 1, 0.f, 0.f, 0);
    }
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

extern "C" int mss_deeplab_anomaly_score_backward(const float *ood_logits, const float *grad_score, int64_t B, int C, int h,
                                                  int w, int H, int W, float *grad_logits, void *stream) {
    MSS_REQUIRE(B >= 0 && C >= 1 && h > 0 && w > 0 && H > 0 && W > 0, "mss_deeplab_anomaly_score_backward: bad shape");
    if (B == 0) return MSS_OK;
    MSS_REQUIRE(ood_logits && grad_score && grad_logits, "mss_deeplab_anomaly_score_backward: null pointer");
    const long long n = (long long)B * h * w, blocks = (n + 255) / 256;
    MSS_REQUIRE(blocks < (1ll << 31), "mss_deeplab_anomaly_score_backward: tensor too large");
    energy_backward_generic_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        ood_logits, grad_score, grad_logits, C, h, w, n, 1, H, W, bw_resize_scale(h, H, 1), bw_resize_scale(w, W, 1), 1);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}
