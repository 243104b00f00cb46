// (Mask2Former half) the mask-logit contraction that sits directly in front of the Mask2Former
// scoring path (mask2former_transformer_decoder.py:528-529 for pred_masks, :548-549 for pred_masks_ood):
//     outputs_mask = torch.einsum("bqc,bchw->bqhw", mask_embed, mask_features)        # [B, Q = 100, h, w]
// mask_features is NCHW fp32 [B, K = 256, h*w]; mask_embed is [B, Q, K] -- per-IMAGE weights.  Per image this is
// the GEMM  D[px, q] = sum_k F[k, px] E[q, k]  with M = pixels, N = Q (padded to 112), K = 256: 51 kFLOP per
// 1.4 KB of traffic -- tensor-core work, HBM-bound.  Kernel: pixel_gemm.cuh (3xTF32 on tcgen05, feature tile
// written from registers straight into TMEM); the pre-split embedding table of an image is 2 x 112 x 256 x 4 B =
// 224 KB, i.e. all the shared memory of the SM's one CTA.  Epilogue: half h of a lane quarter stores query planes
// 56h .. 56h+55 (< Q) in 7 chunks of 8 accumulator columns -- 128-byte coalesced rows per warp.
#include "pixel_gemm.cuh"

namespace mss {

constexpr int MG_N = 112;                          // queries padded: UMMA N % 16 == 0 for M = 128
constexpr int MG_HALF_N = MG_N / 2;                // columns per epilogue half
constexpr int MG_STAGE_K = PG_STAGE_K;
constexpr int MG_MAX_K = PG_MAX_K;

// per image: [hi | lo], each K*MG_N floats; element (k, n) at (k / 4) * (MG_N * 4) + n * 4 + k % 4
__global__ void mask_embed_umma_kernel(const float *__restrict__ embed, int Q, int K, float *__restrict__ table) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // over K * MG_N
    const long long b = blockIdx.y;
    if (i >= K * MG_N) return;
    const int k = i / MG_N, n = i - k * MG_N;
    const float w = (n < Q) ? embed[(b * Q + n) * K + k] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    const int o = (k >> 2) * (MG_N * 4) + n * 4 + (k & 3);
    float *t = table + b * 2 * K * MG_N;
    t[o] = hi;
    t[K * MG_N + o] = w - hi;
}

struct MaskEpi {
    static constexpr int CHUNKS = MG_HALF_N / 8;
    float *out;                                    // [B, Q, hw]
    int Q, hw;
    struct Tile {
        char *o;                                   // next plane to store (null: row past the end)
    };
    __device__ __forceinline__ void begin(Tile &t, long long b, long long p, int half) const {
        t.o = (p >= 0) ? reinterpret_cast<char *>(out + (b * Q + half * MG_HALF_N) * (long long)hw + p) : nullptr;
    }
    __device__ __forceinline__ void chunk(Tile &t, int c, uint32_t d, int half, size_t plane) const {
        uint32_t v[8];
        tc5_ld8(d + half * MG_HALF_N + c * 8, v);
        tc5_wait_ld();
        if (t.o) {
            const int cols = Q - half * MG_HALF_N;  // query planes this thread stores (may be <= 0)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (c * 8 + i < cols) *reinterpret_cast<float *>(t.o) = __uint_as_float(v[i]);
                t.o += plane;
                asm volatile("" : "+l"(t.o));
            }
        }
    }
};

}  // namespace mss

using namespace mss;

extern "C" size_t mss_m2f_mask_logits_workspace_bytes(int64_t B, int K) {
    if (B < 0) B = 0;
    if (K < 0) K = 0;
    return align_up((size_t)B * 2 * K * MG_N * 4, 256) + 512;
}

extern "C" int mss_m2f_mask_logits(const float *mask_embed, const float *mask_features, int64_t B, int Q, int K,
                                   int64_t hw, float *mask_logits, void *workspace, size_t workspace_bytes,
                                   void *stream) {
    MSS_REQUIRE(B >= 0 && hw >= 0 && Q >= 1 && K >= 1, "mss_m2f_mask_logits: bad shape");
    if (B == 0 || hw == 0) return MSS_OK;
    MSS_REQUIRE(mask_embed && mask_features && mask_logits && workspace, "mss_m2f_mask_logits: null pointer");
    if (Q > MG_N || K % MG_STAGE_K != 0 || K > MG_MAX_K || hw > (int64_t)0x7fffffff / MG_MAX_K) {
        set_error("mss_m2f_mask_logits: supported shapes are Q <= %d, K a multiple of %d up to %d, h*w <= %d (got Q=%d K=%d)",
                  MG_N, MG_STAGE_K, MG_MAX_K, 0x7fffffff / MG_MAX_K, Q, K);
        return MSS_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Carver cv(workspace, workspace_bytes);
    float *table = cv.take<float>((size_t)B * 2 * K * MG_N);
    if (!cv.ok()) {
        set_error("mss_m2f_mask_logits: workspace too small (%zu < %zu)", workspace_bytes,
                  mss_m2f_mask_logits_workspace_bytes(B, K));
        return MSS_ERR_WORKSPACE;
    }
    for (int64_t b0 = 0; b0 < B; b0 += 65535) {
        const int nb = (int)std::min<int64_t>(65535, B - b0);
        mask_embed_umma_kernel<<<dim3((K * MG_N + 255) / 256, nb), 256, 0, st>>>(mask_embed + b0 * Q * K, Q, K,
                                                                                  table + b0 * 2 * K * MG_N);
        MSS_CHECK_LAUNCH();
    }
    const int tiles_per_image = (int)((hw + 127) / 128);
    const PixelGemmPlan plan = pixel_gemm_plan(B, tiles_per_image, sm_count());
    const MaskEpi epi{mask_logits, Q, (int)hw};
    const size_t smem = pixel_gemm_smem(K, MG_N);
#define MG_LAUNCH(S)                                                                                                  \
    case S:                                                                                                           \
        MSS_CHECK_CUDA(cudaFuncSetAttribute(pixel_gemm_kernel<S, MG_N, MaskEpi>,                                      \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        pixel_gemm_kernel<S, MG_N, MaskEpi><<<plan.grid, PG_THREADS, smem, st>>>(                                     \
            mask_features, (int)hw, plan.n_items, plan.slices, tiles_per_image, table, (long long)2 * K * MG_N, epi); \
        break;
    switch (K / MG_STAGE_K) {
        MG_LAUNCH(1) MG_LAUNCH(2) MG_LAUNCH(3) MG_LAUNCH(4) MG_LAUNCH(5) MG_LAUNCH(6) MG_LAUNCH(7) MG_LAUNCH(8)
    }
#undef MG_LAUNCH
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}
