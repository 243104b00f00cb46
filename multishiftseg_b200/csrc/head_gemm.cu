//  DeepLabv3+ head fusion: the two bias-free 1x1 convolutions that sit directly in front of the
// scoring path, and the energy score, in ONE pass over the decoder feature map (deepv3.py:279-283):
//     dec1 = self.final[-1](feature)                     # [B, 19, h, w]   class logits   (deepv3.py:279)
//     dec2 = self.ood_head(feature)                      # [B, 19, h, w]   OOD-head logits (deepv3.py:282)
//     energy = -(1. * torch.logsumexp(dec2, dim=1))      # [B, h, w]                       (deepv3.py:251-253)
// feature is NCHW fp32 [B, K = 256, h*w]; it is read once (1 KB per pixel), dec2 never has to exist in HBM.
//
// A 1x1 convolution is the GEMM  D[px, n] = sum_k F[k, px] W[n, k]  with M = pixels, N = 2 x 19 (padded to 48),
// K = 256: 19.5 kFLOP per 1.1 KB of traffic -- above the FP32-FMA ridge of B200 but far below the tensor-core
// one, so it runs on tcgen05 and is HBM-bound.  Kernel: pixel_gemm.cuh (3xTF32, feature tile written from
// registers straight into TMEM, register ring one tile deep); B = the two weight matrices, pre-split, resident in
// shared memory (dec1 at columns 0..23, dec2 at 24..47).  Epilogue (3 chunks of 8 columns, interleaved with the
// next tile's stages): half 0 of a lane quarter stores the dec1 planes; half 1 stores dec2 (if asked) and, on the
// last chunk, re-reads its 24 dec2 columns from TMEM and turns them into the energy with torch's formula.
//
// Round-1 measurements (8 x 256 x 512 x 1024 features, B200):
//   first form (2 CTAs/SM, 2-deep register buffer, predicated 64-bit addressing): 1.20 ms = 3.86 TB/s (59 % of the HBM
//   peak), 1.57x cuDNN's TF32 path for the two convolutions + logsumexp; deeper look-ahead in THAT structure was
//   slower (1.56 ms), an L2 prefetch of the next tile 4 % slower;
//   running-pointer loads + uniform warp index: 0.97 ms = 4.76 TB/s (73 %); then the shared pixel_gemm kernel.
#include "pixel_gemm.cuh"

namespace mss {

constexpr int HG_N = 48, HG_CH = 24;              // N padded; dec1 at columns 0..23, dec2 at 24..47
constexpr int HG_STAGE_K = PG_STAGE_K;
constexpr int HG_MAX_K = PG_MAX_K;

// element (k, n) of a B table: (k / 4) * (HG_N * 4) + n * 4 + k % 4   (8 x 16-byte core matrices, K-major);
// table = [hi | lo], each K * HG_N floats
__global__ void head_weights_umma_kernel(const float *__restrict__ w_cls, const float *__restrict__ w_ood, int C, int K,
                                         float *__restrict__ table) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // over K * HG_N
    if (i >= K * HG_N) return;
    const int k = i / HG_N, n = i - k * HG_N;
    float w = 0.f;
    if (n < C) w = w_cls[(long long)n * K + k];
    else if (n >= HG_CH && n - HG_CH < C) w = w_ood[(long long)(n - HG_CH) * K + k];
    const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    const int o = (k >> 2) * (HG_N * 4) + n * 4 + (k & 3);
    table[o] = hi;
    table[K * HG_N + o] = w - hi;
}

struct HeadEpi {
    static constexpr int CHUNKS = HG_CH / 8;
    float *dec1, *dec2, *energy;                   // [B, C, hw], [B, C, hw], [B, hw]; any may be null
    int C, hw;
    struct Tile {
        long long bp;                              // b * hw + p of this thread's pixel (-1: row past the end)
        int b;
    };
    __device__ __forceinline__ void begin(Tile &t, long long b, long long p, int half) const {
        t.bp = p >= 0 ? b * (long long)hw + p : -1;
        t.b = (int)b;
    }
    __device__ __forceinline__ void chunk(Tile &t, int c, uint32_t d, int half, size_t plane) const {
        const uint32_t col = d + half * HG_CH;
        float *planes = half == 0 ? dec1 : dec2;
        if (planes) {                              // (warp-uniform: tcgen05.ld is a warp-collective)
            uint32_t v[8];
            tc5_ld8(col + c * 8, v);
            tc5_wait_ld();
            if (t.bp >= 0) {
                // plane c*8 of image b: planes + (b*C + c*8)*hw + p
                char *o = reinterpret_cast<char *>(planes + t.bp + ((long long)t.b * (C - 1) + c * 8) * (long long)hw);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (c * 8 + i < C) stg_stream_f1(reinterpret_cast<float *>(o), __uint_as_float(v[i]));
                    o += plane;
                    asm volatile("" : "+l"(o));
                }
            }
        }
        if (c == CHUNKS - 1 && half == 1 && energy) {
            // torch.logsumexp: m = amax; m' = isinf(m) ? 0 : m; log(sum exp(x - m')) + m'.  Two sweeps over the 24
            // dec2 columns in TMEM (reads are cheap; 24 live registers next to the load ring are not)
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < HG_CH; k += 8) {
                uint32_t v[8];
                tc5_ld8(col + k, v);
                tc5_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (k + j < C) mx = fmaxf(mx, __uint_as_float(v[j]));
            }
            const float ms = isinf(mx) ? 0.f : mx;
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < HG_CH; k += 8) {
                uint32_t v[8];
                tc5_ld8(col + k, v);
                tc5_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (k + j < C) sum += expf(__uint_as_float(v[j]) - ms);
            }
            if (t.bp >= 0) stg_stream_f1(energy + t.bp, -(logf(sum) + ms));
        }
    }
};

}  // namespace mss

using namespace mss;

extern "C" size_t mss_deeplab_head_workspace_bytes(int K) {
    if (K < 0) K = 0;
    return align_up((size_t)2 * K * HG_N * 4, 256) + 512;
}

extern "C" int mss_deeplab_head(const float *feature, int64_t B, int K, int64_t hw, const float *w_cls,
                                const float *w_ood, int C, float *dec1, float *dec2, float *energy, void *workspace,
                                size_t workspace_bytes, void *stream) {
    MSS_REQUIRE(B >= 0 && hw >= 0 && C >= 1 && K >= 1, "mss_deeplab_head: bad shape");
    MSS_REQUIRE(dec1 || dec2 || energy, "mss_deeplab_head: no output requested");
    if (B == 0 || hw == 0) return MSS_OK;
    MSS_REQUIRE(feature && w_cls && w_ood && workspace, "mss_deeplab_head: null pointer");
    if (C > HG_CH || K % HG_STAGE_K != 0 || K > HG_MAX_K || hw > (int64_t)0x7fffffff / HG_MAX_K) {
        set_error("mss_deeplab_head: supported shapes are C <= %d, K a multiple of %d up to %d, h*w <= %d (got C=%d K=%d)",
                  HG_CH, HG_STAGE_K, HG_MAX_K, 0x7fffffff / HG_MAX_K, C, K);
        return MSS_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Carver cv(workspace, workspace_bytes);
    float *table = cv.take<float>((size_t)2 * K * HG_N);
    if (!cv.ok()) {
        set_error("mss_deeplab_head: workspace too small (%zu < %zu)", workspace_bytes, mss_deeplab_head_workspace_bytes(K));
        return MSS_ERR_WORKSPACE;
    }
    head_weights_umma_kernel<<<(K * HG_N + 255) / 256, 256, 0, st>>>(w_cls, w_ood, C, K, table);
    MSS_CHECK_LAUNCH();
    const int tiles_per_image = (int)((hw + 127) / 128);
    const PixelGemmPlan plan = pixel_gemm_plan(B, tiles_per_image, sm_count(), /*shared_table=*/true);
    const HeadEpi epi{dec1, dec2, energy, C, (int)hw};
    const size_t smem = pixel_gemm_smem(K, HG_N);
#define HG_LAUNCH(S)                                                                                                  \
    case S:                                                                                                           \
        MSS_CHECK_CUDA(cudaFuncSetAttribute(pixel_gemm_kernel<S, HG_N, HeadEpi>,                                      \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        pixel_gemm_kernel<S, HG_N, HeadEpi><<<plan.grid, PG_THREADS, smem, st>>>(                                     \
            feature, (int)hw, plan.n_items, plan.slices, tiles_per_image, table, 0ll, epi);                           \
        break;
    switch (K / HG_STAGE_K) {
        HG_LAUNCH(1) HG_LAUNCH(2) HG_LAUNCH(3) HG_LAUNCH(4) HG_LAUNCH(5) HG_LAUNCH(6) HG_LAUNCH(7) HG_LAUNCH(8)
    }
#undef HG_LAUNCH
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}
