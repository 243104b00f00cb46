// (Mask2Former half) the mask-logit contraction that sits directly in front of the Mask2Former
// scoring path (mask2former_transformer_decoder.py:528-529 for pred_masks, :548-549 for pred_masks_ood):
//     outputs_mask = torch.einsum("bqc,bchw->bqhw", mask_embed, mask_features)        # [B, Q = 100, h, w]
// mask_features is NCHW fp32 [B, K = 256, h*w]; mask_embed is [B, Q, K] -- per-IMAGE weights.  Per image this is
// the GEMM  D[px, q] = sum_k F[k, px] E[q, k]  with M = pixels, N = Q (padded to 112), K = 256: 51 kFLOP per
// 1.4 KB of traffic -- tensor-core work, HBM-bound.  Same numerics and operand flow as head_gemm.cu:
//     F = F_hi + F_lo, E = E_hi + E_lo (hi = top 19 bits, lo = exact remainder)
//     D = F_lo*E_hi + F_hi*E_lo + F_hi*E_hi            (3xTF32, fp32 accumulation in TMEM)
// with the feature values loaded coalesced (lane <-> pixel <-> TMEM lane), split in registers and written
// straight into TENSOR MEMORY as the A operand; the pre-split embedding table of the image is the B operand.
//
// What differs from the DeepLab head: the B table is 2 x 112 x 256 x 4 B = 224 KB -- one CTA per SM, all of its
// shared memory -- and it changes with the image.  So a CTA is TWO independent pipelines ("groups", each the
// producer / issuer / epilogue structure of head_gemm.cu on its own 128-pixel tiles) sharing one B table:
//   * warps 0-7 / 8-15: producers + epilogue of group 0 / 1; warp 16 / 17: MMA issuer of group 0 / 1;
//   * work item = (image, slice): a CTA keeps one image's table and walks tiles slice, slice + S, ... of that
//     image (S = SMs / B slices per image when B <= SMs: one item per CTA, ~1000/S tiles each);
//   * the table arrives as ONE 224 KB bulk copy (cp.async.bulk -> mbarrier), issued by the MMA warps once both
//     groups' tensor work of the previous item has completed;
//   * TMEM (all 512 columns): accumulators D_g at g*128 (112 columns), A buffers at 256 + g*128 + slot*64.
//     One accumulator per group: the epilogue of a tile runs before the group's next tile is staged (the other
//     group keeps the memory system busy meanwhile); the next tile's first 16 loads are already in flight.
//   * epilogue: half h of a lane quarter stores query planes 56h .. 56h+55 (< Q): 128-byte coalesced rows.
#include "tc5_common.cuh"

namespace mss {

constexpr int MG_N = 112;                          // queries padded: UMMA N % 16 == 0 for M = 128
constexpr int MG_HALF_N = MG_N / 2;                // columns per epilogue half
constexpr int MG_STAGE_K = 32;
constexpr int MG_GROUP_THREADS = 256;              // producer threads per group
constexpr int MG_THREADS = 2 * MG_GROUP_THREADS + 64;
constexpr int MG_TMEM_COLS = 512;
constexpr int MG_MAX_K = 256;
constexpr uint32_t MG_IDESC = tc5_idesc_tf32(128, MG_N);

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

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__global__ void __launch_bounds__(MG_THREADS, 1)
mask_gemm_kernel(const float *__restrict__ feat, int hw, int K, int Q, long long n_items, int slices,
                 int tiles_per_image, const float *__restrict__ table, float *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_b = reinterpret_cast<float *>(smem_raw);                            // [hi | lo][K/4][112][4]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_b + 2 * K * MG_N);          // per group: full[2] empty[2] dfull dempty; + table
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 16);
    uint64_t *bar_table = s_bar + 12;

    const int tid = threadIdx.x, lane = tid & 31;
    // warp index through a shuffle: ptxas then knows it is warp-uniform, keeps the role branches uniform (BRA.U) and
    // the load descriptors / loop state in uniform registers instead of re-materialising them (R2UR) at every load
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int stages = K / MG_STAGE_K;
    const unsigned table_bytes = (unsigned)(2 * K * MG_N * 4);

    if (tid == 2 * MG_GROUP_THREADS) {
        for (int g = 0; g < 2; g++) {
            uint64_t *bg = s_bar + 6 * g;
            mbar_init(&bg[0], MG_GROUP_THREADS);
            mbar_init(&bg[1], MG_GROUP_THREADS);
            mbar_init(&bg[2], 1);
            mbar_init(&bg[3], 1);
            mbar_init(&bg[4], 1);
            mbar_init(&bg[5], MG_GROUP_THREADS);
        }
        mbar_init(bar_table, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "n"(MG_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *s_tmem;

    if (warp < 16) {
        // ===== producers + epilogue of group g =====
        const int g = warp >> 3, wg = warp & 7;
        const int quarter = wg & 3, half = wg >> 2;
        const int m = quarter * 32 + lane;                                       // pixel inside the tile == TMEM lane
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
        const uint32_t col_d = (uint32_t)g * 128, col_a = 256 + (uint32_t)g * 128;
        uint64_t *bar_full = s_bar + 6 * g, *bar_empty = bar_full + 2, *bar_dfull = bar_full + 4, *bar_dempty = bar_full + 5;

        unsigned u = 0, done = 0;                                                // stage uses / finished epilogues (whole kernel; only parities matter)
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
            const long long b = item / slices;
            const int slice = (int)(item - b * slices);
            const int n_j = (tiles_per_image > slice) ? (tiles_per_image - slice + slices - 1) / slices : 0;
            // offsets inside one image fit 32 bits (host-checked: K * hw, Q * hw < 2^31): fewer 64-bit registers
            const float *fimg = feat + (b * K + half * 16) * (long long)hw;
            float *oimg = out + (b * Q + half * MG_HALF_N) * (long long)hw;

            // epilogue of the tile whose first pixel row is `p`: columns 56*half .. 56*half + 55 of this lane's row
            auto epilogue = [&](int p) {
                mbar_wait(bar_dfull, (unsigned)(done & 1));
                tc5_fence_after();
                const uint32_t d = lane_base + col_d + half * MG_HALF_N;
                float *o = oimg + (p < hw ? p : 0);
                const int ncols = (p < hw) ? min(MG_HALF_N, Q - half * MG_HALF_N) : 0;     // query planes this thread stores
                const size_t plane = (size_t)(unsigned)hw * sizeof(float);
#pragma unroll
                for (int c0 = 0; c0 < MG_HALF_N; c0 += 8) {
                    uint32_t v[8];
                    tc5_ld8(d + c0, v);
                    tc5_wait_ld();
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        if (c0 + c < ncols) *o = __uint_as_float(v[c]);
                        o = reinterpret_cast<float *>(reinterpret_cast<char *>(o) + plane);
                    }
                }
                tc5_fence_before();
                mbar_arrive(bar_dempty);                                         // the accumulator may be overwritten
                done++;
            };

            int p_prev = -1;
            for (int j = g; j < n_j; j += 2) {
                const long long p64 = ((long long)slice + (long long)j * slices) * 128 + m;
                const int p = (int)(p64 < hw ? p64 : hw);   // hw = "past the end": loads are clamped, stores skipped
                // One running pointer per tile, advanced by a plane per load: 2 instructions per element.  (First form:
                // `live ? ld(src + i * hw) : 0` -- ncu/SASS showed ~8 instructions of predicated 64-bit address
                // arithmetic per load, 21 warp instructions per element in total, issue slots 50 % busy with 4.5 warps
                // per scheduler: the kernel was issue/latency-bound at 4.0 TB/s.)
                const float *q = fimg + (p < hw ? p : hw - 1);
                const size_t plane = (size_t)(unsigned)hw * sizeof(float);
                float cur[16], nxt[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    cur[i] = ldg_stream_f1(q);
                    q = reinterpret_cast<const float *>(reinterpret_cast<const char *>(q) + plane);
                }
                // the previous tile's epilogue runs here, behind this tile's first 16 loads and before the stage-1 loads
                // are issued: no second register buffer is live across it (inside the stage loop it spilled `nxt`)
                if (p_prev >= 0) epilogue(p_prev);
                for (int s = 0; s < stages; s++, u++) {
                    if (s + 1 < stages) {
                        q = reinterpret_cast<const float *>(reinterpret_cast<const char *>(q) + 16 * plane);   // the other half's channels
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            nxt[i] = ldg_stream_f1(q);
                            q = reinterpret_cast<const float *>(reinterpret_cast<const char *>(q) + plane);
                        }
                    }
                    const int slot = (int)(u & 1);
                    if (u >= 2) mbar_wait(&bar_empty[slot], (unsigned)(((u >> 1) + 1) & 1));   // MMAs of use u-2 are done
                    tc5_fence_after();
                    const uint32_t a = lane_base + col_a + slot * 64 + half * 16;
#pragma unroll
                    for (int c = 0; c < 2; c++) {                                // 8 channels at a time: 16 temporaries, not 32
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
#pragma unroll
                    for (int i = 0; i < 16; i++) cur[i] = nxt[i];
                }
                p_prev = p;
            }
            if (p_prev >= 0) epilogue(p_prev);
        }
    } else {
        // ===== MMA issuer of group g: the whole warp waits (stays converged), one elected lane issues =====
        const int g = warp - 16;
        uint64_t *bar_full = s_bar + 6 * g, *bar_empty = bar_full + 2, *bar_dfull = bar_full + 4, *bar_dempty = bar_full + 5;
        const uint32_t bhi = smem_u32(s_b), blo = bhi + (uint32_t)K * MG_N * 4;
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);               // warp-uniform (see m2f_tc5q.cuh)
        const uint32_t d = tmem_u + (uint32_t)g * 128, col_a = tmem_u + 256 + (uint32_t)g * 128;
        unsigned u = 0, done = 0, n_loaded = 0;
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x, n_loaded++) {
            const long long b = item / slices;
            const int slice = (int)(item - b * slices);
            const int n_j = (tiles_per_image > slice) ? (tiles_per_image - slice + slices - 1) / slices : 0;
            // the image's table: every MMA of both groups that read the previous one has completed
            if (n_loaded > 0) {
                if (done > 0) mbar_wait_backoff(bar_dfull, (unsigned)((done - 1) & 1), 32);
                named_bar_sync(1, 64);
            }
            if (g == 0 && elect_one_sync()) {
                mbar_expect_tx(bar_table, table_bytes);
                bulk_load_1d(s_b, table + b * 2 * K * MG_N, table_bytes, bar_table);
            }
            __syncwarp();
            mbar_wait_backoff(bar_table, (unsigned)(n_loaded & 1), 32);
            for (int j = g; j < n_j; j += 2, done++) {
                for (int s = 0; s < stages; s++, u++) {
                    const int slot = (int)(u & 1);
                    mbar_wait_backoff(&bar_full[slot], (unsigned)((u >> 1) & 1), 32);
                    if (s == 0 && done > 0) mbar_wait_backoff(bar_dempty, (unsigned)((done - 1) & 1), 32);   // epilogue has read the tile before
                    tc5_fence_after();
                    if (elect_one_sync()) {
#pragma unroll
                        for (int kk = 0; kk < MG_STAGE_K / 8; kk++) {
                            const int ks = s * (MG_STAGE_K / 8) + kk;            // k-step of 8 channels = 2 core-matrix chunks
                            const uint64_t dh = tc5_smem_desc(bhi + ks * 2 * (MG_N * 16), MG_N * 16, 128);
                            const uint64_t dl = tc5_smem_desc(blo + ks * 2 * (MG_N * 16), MG_N * 16, 128);
                            const uint32_t ahi = col_a + slot * 64 + kk * 8, alo = ahi + 32;
                            tc5_mma_ts(d, alo, dh, MG_IDESC, (s | kk) > 0);
                            tc5_mma_ts(d, ahi, dl, MG_IDESC, 1);
                            tc5_mma_ts(d, ahi, dh, MG_IDESC, 1);
                        }
                        tc5_commit(&bar_empty[slot]);
                        if (s == stages - 1) tc5_commit(bar_dfull);
                    }
                    __syncwarp();
                }
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc5_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(MG_TMEM_COLS) : "memory");
    }
}

static size_t mask_gemm_smem(int K) { return (size_t)2 * K * MG_N * 4 + 16 * 8 + 16; }

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
    static std::atomic<bool> attr_set{false};
    if (!attr_set.load()) {
        MSS_CHECK_CUDA(cudaFuncSetAttribute(mask_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)mask_gemm_smem(MG_MAX_K)));
        attr_set.store(true);
    }
    const int tiles_per_image = (int)((hw + 127) / 128);
    const int sms = sm_count();
    // B <= SMs: S = SMs / B slices per image, one work item per CTA; otherwise whole images, round-robin
    int slices = (B <= sms) ? std::min(tiles_per_image, sms / (int)B) : 1;
    if (slices < 1) slices = 1;
    const long long n_items = (long long)B * slices;
    const int grid = (int)std::min<long long>(n_items, (long long)sms);
    mask_gemm_kernel<<<grid, MG_THREADS, mask_gemm_smem(K), st>>>(mask_features, (int)hw, K, Q, n_items, slices, tiles_per_image,
                                                                 table, mask_logits);
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}
