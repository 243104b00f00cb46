//  DeepLabv3+ head fusion: the two bias-free 1x1 convolutions that sit directly in front of the
// scoring path, and the energy score, in ONE pass over the decoder feature map (deepv3.py:279-283):
//     dec1 = self.final[-1](feature)                     # [B, 19, h, w]   class logits   (deepv3.py:279)
//     dec2 = self.ood_head(feature)                      # [B, 19, h, w]   OOD-head logits (deepv3.py:282)
//     energy = -(1. * torch.logsumexp(dec2, dim=1))      # [B, h, w]                       (deepv3.py:251-253)
// feature is NCHW fp32 [B, K = 256, h*w]; it is read once (1 KB per pixel), dec2 never has to exist in HBM.
//
// A 1x1 convolution is the GEMM  D[px, n] = sum_k F[k, px] W[n, k]  with M = pixels, N = 2 x 19 (padded to 48),
// K = 256: 19.5 kFLOP per 1.1 KB of traffic -- above the FP32-FMA ridge of B200 but far below the tensor-core
// one, so it runs on tcgen05 and is HBM-bound.  3xTF32 keeps fp32-level accuracy:
//     F = F_hi + F_lo, W = W_hi + W_lo (hi = top 19 bits, lo = exact remainder; the tensor core reads 19 bits)
//     D = F_lo*W_hi + F_hi*W_lo + F_hi*W_hi            (fp32 accumulation in TMEM)
// The feature map is pixel-contiguous (NCHW), i.e. "M-major"; instead of staging it through a swizzled shared
// memory layout, the producer warps read it with plain coalesced loads (lane <-> pixel <-> TMEM lane), split it
// in registers and write it straight into TENSOR MEMORY as the A operand (tcgen05.st), exactly like the
// Mask2Former kernel does with its sigmoids.
//   * persistent CTA (2 per SM), tile = 128 consecutive pixels of one image;
//   * warps 0-7 (producers): warp = (lane quarter, half); stage = 32 channels, half h takes 16 of them:
//     16 independent 128-byte-per-warp loads (next stage prefetched in registers), split, 2 x tcgen05.st.x16;
//     two A buffers alternate;
//   * warp 8, one elected lane: per stage 4 k-steps x 3 tcgen05.mma (M = 128, N = 48, K = 8), B = the two
//     weight matrices, pre-split, in shared memory for the whole kernel (K-major core matrices, no swizzle);
//   * epilogue one tile behind (two accumulator buffers): half 0 stores the 19 dec1 planes, half 1 turns the 19
//     dec2 columns into the energy (and stores dec2 if asked) -- 128-byte coalesced rows.
// TMEM: D buffers at columns 0 and 64, A buffers at 128 and 192 (hi at +0..31, lo at +32..63) -> 256 columns.
//
// Round-1 measurements (8 x 256 x 512 x 1024 features, B200): 1.20 ms = 3.86 TB/s of algorithmic traffic (59 % of the
// HBM peak, DRAM reads = 1x the feature bytes), 1.57x faster than cuDNN's TF32 path for the two convolutions +
// logsumexp.  ncu: 25 % of the stall samples wait for the prefetched feature registers.  Deeper look-ahead made
// it SLOWER (1.56 ms with 16-channel stages issued 3 stages ahead, with or without L1 allocation, cyclic or
// blocked tile order): the observed load latency grows with the bytes in flight (1.3 us -> 2.6 us), i.e. the
// memory side saturates near 3-4 TB/s for this access pattern -- 256 planes of 2 MB, every 128-byte request of a
// warp on a different page -- rather than the SM running out of requests.  Next: feed the tile through TMA
// (2-D box over [channels x pixels]) so a tile is a few large requests instead of 256 x 4 small ones.
#include "tc5_common.cuh"

namespace mss {

constexpr int HG_N = 48, HG_CH = 24;              // N padded; dec1 at columns 0..23, dec2 at 24..47
constexpr int HG_STAGE_K = 32;
constexpr int HG_THREADS = 288, HG_PRODUCERS = 256;
constexpr int HG_TMEM_COLS = 256;
constexpr int HG_COL_D = 0, HG_COL_A = 128;
constexpr int HG_MAX_K = 256;
constexpr uint32_t HG_IDESC = tc5_idesc_tf32(128, HG_N);

// element (k, n) of a B table: (k / 4) * (HG_N * 4) + n * 4 + k % 4   (8 x 16-byte core matrices, K-major)
__global__ void head_weights_umma_kernel(const float *__restrict__ w_cls, const float *__restrict__ w_ood, int C, int K,
                                         float *__restrict__ b_hi, float *__restrict__ b_lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // over K * HG_N
    if (i >= K * HG_N) return;
    const int k = i / HG_N, n = i - k * HG_N;
    float w = 0.f;
    if (n < C) w = w_cls[(long long)n * K + k];
    else if (n >= HG_CH && n - HG_CH < C) w = w_ood[(long long)(n - HG_CH) * K + k];
    const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    const int o = (k >> 2) * (HG_N * 4) + n * 4 + (k & 3);
    b_hi[o] = hi;
    b_lo[o] = w - hi;
}

struct HeadOut {
    float *dec1, *dec2, *energy;
};

__global__ void __launch_bounds__(HG_THREADS, 2)
head_gemm_kernel(const float *__restrict__ feat, long long hw, int K, int C, long long n_tiles, long long tiles_per_image,
                 const float *__restrict__ b_hi, const float *__restrict__ b_lo, HeadOut out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_bhi = reinterpret_cast<float *>(smem_raw);                         // [K/4][48][4]
    float *s_blo = s_bhi + K * HG_N;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_blo + K * HG_N);           // full[2] empty[2] dfull[2]
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 8);
    uint64_t *bar_full = s_bar, *bar_empty = s_bar + 2, *bar_dfull = s_bar + 4;

    const int tid = threadIdx.x, lane = tid & 31;
    // warp index through a shuffle: ptxas then knows it is warp-uniform, keeps the role branches uniform (BRA.U) and
    // the load descriptors / loop state in uniform registers instead of re-materialising them (R2UR) at every load
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int stages = K / HG_STAGE_K;

    if (tid == HG_PRODUCERS) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&bar_full[s], HG_PRODUCERS);
            mbar_init(&bar_empty[s], 1);
            mbar_init(&bar_dfull[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "n"(HG_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the pre-split weights (already in core-matrix order) stay in shared memory for the whole kernel
    for (int i = tid; i < K * HG_N / 4; i += HG_THREADS) {
        reinterpret_cast<float4 *>(s_bhi)[i] = __ldg(reinterpret_cast<const float4 *>(b_hi) + i);
        reinterpret_cast<float4 *>(s_blo)[i] = __ldg(reinterpret_cast<const float4 *>(b_lo) + i);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");               // the tensor core reads them
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *s_tmem;

    const long long my_tiles = (n_tiles > (long long)blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp < 8) {
        // ===== producers + epilogue =====
        const int quarter = warp & 3, half = warp >> 2;
        const int m = quarter * 32 + lane;                                       // pixel inside the tile == TMEM lane
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);

        auto epilogue = [&](long long it) {                                      // `it`-th tile of this CTA
            const long long tile = (long long)blockIdx.x + it * gridDim.x;
            const long long b = tile / tiles_per_image, p = (tile - b * tiles_per_image) * 128 + m;
            const int buf = (int)(it & 1);
            mbar_wait(&bar_dfull[buf], (unsigned)((it >> 1) & 1));
            tc5_fence_after();
            uint32_t v[HG_CH];
            const uint32_t d = lane_base + HG_COL_D + buf * 64 + half * HG_CH;
            {
                uint32_t a[8], bb[8], c[8];
                tc5_ld8(d, a);
                tc5_ld8(d + 8, bb);
                tc5_ld8(d + 16, c);
                tc5_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; j++) { v[j] = a[j]; v[8 + j] = bb[j]; v[16 + j] = c[j]; }
            }
            if (p >= hw) return;
            if (half == 0) {
                if (out.dec1) {
                    float *o = out.dec1 + (b * C) * hw + p;
#pragma unroll
                    for (int c = 0; c < HG_CH; c++)
                        if (c < C) stg_stream_f1(o + (long long)c * hw, __uint_as_float(v[c]));
                }
            } else {
                if (out.dec2) {
                    float *o = out.dec2 + (b * C) * hw + p;
#pragma unroll
                    for (int c = 0; c < HG_CH; c++)
                        if (c < C) stg_stream_f1(o + (long long)c * hw, __uint_as_float(v[c]));
                }
                if (out.energy) {
                    // torch.logsumexp: m = amax; m' = isinf(m) ? 0 : m; log(sum exp(x - m')) + m'
                    float mx = __uint_as_float(v[0]);
#pragma unroll
                    for (int c = 1; c < HG_CH; c++)
                        if (c < C) mx = fmaxf(mx, __uint_as_float(v[c]));
                    const float ms = isinf(mx) ? 0.f : mx;
                    float s = 0.f;
#pragma unroll
                    for (int c = 0; c < HG_CH; c++)
                        if (c < C) s += expf(__uint_as_float(v[c]) - ms);
                    stg_stream_f1(out.energy + b * hw + p, -(logf(s) + ms));
                }
            }
        };

        long long u = 0;                                                         // stage uses so far (A buffer = u & 1)
        for (long long it = 0; it < my_tiles; it++) {
            const long long tile = (long long)blockIdx.x + it * gridDim.x;
            const long long b = tile / tiles_per_image, p = (tile - b * tiles_per_image) * 128 + m;
            // One running pointer per tile, advanced by a plane per load (rows past the end read the last pixel; their
            // results are never stored).  The first form, `live ? ld(src + j * hw) : 0`, cost ~8 instructions of predicated
            // 64-bit address arithmetic per load (SASS, round 1).
            const char *q = reinterpret_cast<const char *>(feat + (b * K + half * 16) * hw + (p < hw ? p : hw - 1));
            const size_t plane = (size_t)hw * sizeof(float);
            float cur[16], nxt[16];
#pragma unroll
            for (int j = 0; j < 16; j++) { cur[j] = ldg_stream_f1(reinterpret_cast<const float *>(q)); q += plane; }
            for (int s = 0; s < stages; s++, u++) {
                if (s + 1 < stages) {
                    q += 16 * plane;                                             // the other half's channels
#pragma unroll
                    for (int j = 0; j < 16; j++) { nxt[j] = ldg_stream_f1(reinterpret_cast<const float *>(q)); q += plane; }
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    hi[j] = __float_as_uint(cur[j]) & 0xFFFFE000u;
                    lo[j] = __float_as_uint(cur[j] - __uint_as_float(hi[j]));
                }
                const int slot = (int)(u & 1);
                if (u >= 2) mbar_wait(&bar_empty[slot], (unsigned)(((u >> 1) + 1) & 1));   // MMAs of use u-2 are done
                tc5_fence_after();
                const uint32_t a = lane_base + HG_COL_A + slot * 64 + half * 16;
                tc5_st16(a, hi);
                tc5_st16(a + 32, lo);
                tc5_wait_st();
                tc5_fence_before();
                mbar_arrive(&bar_full[slot]);
                if (s == 0 && it > 0) epilogue(it - 1);                          // one tile behind: its MMAs are long done
#pragma unroll
                for (int j = 0; j < 16; j++) cur[j] = nxt[j];
            }
        }
        if (my_tiles > 0) epilogue(my_tiles - 1);
    } else {
        // ===== MMA issuer: the whole warp waits (stays converged), one elected lane issues =====
        const uint32_t bhi = smem_u32(s_bhi), blo = smem_u32(s_blo);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);               // warp-uniform (see m2f_tc5q.cuh)
        long long u = 0;
        for (long long it = 0; it < my_tiles; it++) {
            const uint32_t d = tmem_u + HG_COL_D + (uint32_t)(it & 1) * 64;
            for (int s = 0; s < stages; s++, u++) {
                const int slot = (int)(u & 1);
                mbar_wait_backoff(&bar_full[slot], (unsigned)((u >> 1) & 1), 32);
                tc5_fence_after();
                if (elect_one_sync()) {
#pragma unroll
                    for (int kk = 0; kk < HG_STAGE_K / 8; kk++) {
                        const int ks = s * (HG_STAGE_K / 8) + kk;                // k-step of 8 channels = 2 core-matrix chunks
                        const uint64_t dh = tc5_smem_desc(bhi + ks * 2 * (HG_N * 16), HG_N * 16, 128);
                        const uint64_t dl = tc5_smem_desc(blo + ks * 2 * (HG_N * 16), HG_N * 16, 128);
                        const uint32_t ahi = tmem_u + HG_COL_A + slot * 64 + kk * 8, alo = ahi + 32;
                        tc5_mma_ts(d, alo, dh, HG_IDESC, (s | kk) > 0);
                        tc5_mma_ts(d, ahi, dl, HG_IDESC, 1);
                        tc5_mma_ts(d, ahi, dh, HG_IDESC, 1);
                    }
                    tc5_commit(&bar_empty[slot]);
                    if (s == stages - 1) tc5_commit(&bar_dfull[it & 1]);
                }
                __syncwarp();
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc5_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(HG_TMEM_COLS) : "memory");
    }
}

static size_t head_smem(int K) { return (size_t)2 * K * HG_N * 4 + 8 * 8 + 16 + 128; }

}  // namespace mss

using namespace mss;

extern "C" size_t mss_deeplab_head_workspace_bytes(int K) {
    if (K < 0) K = 0;
    return 2 * align_up((size_t)K * HG_N * 4, 256) + 512;
}

extern "C" int mss_deeplab_head(const float *feature, int64_t B, int K, int64_t hw, const float *w_cls,
                                const float *w_ood, int C, float *dec1, float *dec2, float *energy, void *workspace,
                                size_t workspace_bytes, void *stream) {
    MSS_REQUIRE(B >= 0 && hw >= 0 && C >= 1 && K >= 1, "mss_deeplab_head: bad shape");
    MSS_REQUIRE(dec1 || dec2 || energy, "mss_deeplab_head: no output requested");
    if (B == 0 || hw == 0) return MSS_OK;
    MSS_REQUIRE(feature && w_cls && w_ood && workspace, "mss_deeplab_head: null pointer");
    if (C > HG_CH || K % HG_STAGE_K != 0 || K > HG_MAX_K) {
        set_error("mss_deeplab_head: supported shapes are C <= %d, K a multiple of %d up to %d (got C=%d K=%d)", HG_CH,
                  HG_STAGE_K, HG_MAX_K, C, K);
        return MSS_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Carver cv(workspace, workspace_bytes);
    float *b_hi = cv.take<float>((size_t)K * HG_N);
    float *b_lo = cv.take<float>((size_t)K * HG_N);
    if (!cv.ok()) {
        set_error("mss_deeplab_head: workspace too small (%zu < %zu)", workspace_bytes, mss_deeplab_head_workspace_bytes(K));
        return MSS_ERR_WORKSPACE;
    }
    head_weights_umma_kernel<<<(K * HG_N + 255) / 256, 256, 0, st>>>(w_cls, w_ood, C, K, b_hi, b_lo);
    MSS_CHECK_LAUNCH();
    const size_t smem = head_smem(K);
    static std::atomic<bool> attr_set{false};
    if (!attr_set.load()) {
        MSS_CHECK_CUDA(cudaFuncSetAttribute(head_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_smem(HG_MAX_K)));
        attr_set.store(true);
    }
    const long long tiles_per_image = (hw + 127) / 128, n_tiles = tiles_per_image * B;
    const int grid = (int)std::min<long long>(n_tiles, (long long)sm_count() * 2);
    head_gemm_kernel<<<grid, HG_THREADS, smem, st>>>(feature, hw, K, C, n_tiles, tiles_per_image, b_hi, b_lo,
                                                     HeadOut{dec1, dec2, energy});
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}
