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
// The B table is 2 x 112 x 256 x 4 B = 224 KB -- one CTA per SM, all of its shared memory, nothing left to stage
// features in -- and it changes with the image.  The feature tile therefore waits in REGISTERS:
//   * 8 producer warps + 1 MMA-issuer warp (288 threads, ~200 registers each); tile = 128 pixels, stage = 32
//     channels (a thread's half takes 16 of them);
//   * a thread keeps a ring of 6 x 16 loaded values (96 registers; three quarters of a K = 256 tile): right after a
//     stage has been split and stored to TMEM, the loads of the stage 6 positions later -- the rest of this tile, then
//     the next one -- are issued into the same registers, so every load has ~6 us to arrive and ~90 KB per SM are
//     in flight.
//     (v1 of this kernel, two 8-warp pipelines with a 2-deep register buffer at 96 registers/thread: 0.357 ms for the
//     cfg-3 batch, 41 % of all stall samples on the first use of a loaded value, DRAM 47 %, tensor pipe 50 % -- the
//     load latency and the tensor work added up instead of overlapping.)
//   * 4 A slots in TMEM (64 columns each: 32 hi + 32 lo), two accumulators (112 columns each) -> all 512 columns;
//   * work item = (image, slice): a CTA keeps one image's table and walks tiles slice, slice + S, ... of that
//     image (S = SMs / B slices per image when B <= SMs: one item per CTA); the table arrives as ONE 224 KB bulk
//     copy (cp.async.bulk -> mbarrier) issued by the MMA warp once the tensor work of the previous item is complete;
//     the producers never touch it and prefetch straight across item boundaries;
//   * epilogue of tile t: 7 chunks of 8 accumulator columns, one chunk after each stage of tile t+1 (half h of a lane
//     quarter stores query planes 56h .. 56h+55 (< Q): 128-byte coalesced rows).
// Addresses: one running pointer per tile, advanced by a plane per load (2 instructions per element); warp index
// through a shuffle so that ptxas keeps the role branches and descriptors uniform.
#include <type_traits>

#include "tc5_common.cuh"

namespace mss {

constexpr int MG_N = 112;                          // queries padded: UMMA N % 16 == 0 for M = 128
constexpr int MG_HALF_N = MG_N / 2;                // columns per epilogue half
constexpr int MG_CHUNKS = MG_HALF_N / 8;           // epilogue chunks of 8 columns
constexpr int MG_STAGE_K = 32;
constexpr int MG_PRODUCERS = 256;
constexpr int MG_THREADS = MG_PRODUCERS + 32;
constexpr int MG_SLOTS = 4;
constexpr int MG_TMEM_COLS = 512;
constexpr int MG_COL_A = 256;                      // D buffers at 0 and 128, A slot k at 256 + 64 k
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

// the CTA's flat tile sequence: items blockIdx.x, + gridDim.x, ...; inside item (image b, slice): tiles slice + j * slices
struct TileCursor {
    long long item, n_items, b;
    int slices, tiles_per_image, slice, j, n_j;
    __device__ __forceinline__ void set_item() {
        b = item / slices;
        slice = (int)(item - b * slices);
        n_j = (tiles_per_image - slice + slices - 1) / slices;      // >= 1: the host keeps slices <= tiles_per_image
        j = 0;
    }
    __device__ __forceinline__ TileCursor(long long first, long long n_items_, int slices_, int tpi)
        : item(first), n_items(n_items_), b(0), slices(slices_), tiles_per_image(tpi), slice(0), j(0), n_j(0) {
        if (item < n_items) set_item();
    }
    __device__ __forceinline__ bool valid() const { return item < n_items; }
    __device__ __forceinline__ long long first_pixel() const { return ((long long)slice + (long long)j * slices) * 128; }
    __device__ __forceinline__ void next() {
        if (++j >= n_j) {
            item += gridDim.x;
            if (item < n_items) set_item();
        }
    }
};

template <int STAGES>
__global__ void __launch_bounds__(MG_THREADS, 1)
mask_gemm_kernel(const float *__restrict__ feat, int hw, int Q, long long n_items, int slices, int tiles_per_image,
                 const float *__restrict__ table, float *__restrict__ out) {
    constexpr int K = STAGES * MG_STAGE_K;
    constexpr int CPS = (MG_CHUNKS + STAGES - 1) / STAGES;                       // epilogue chunks per stage
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_b = reinterpret_cast<float *>(smem_raw);                            // [hi | lo][K/4][112][4]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_b + 2 * K * MG_N);          // full[4] empty[4] dfull[2] dempty[2] table
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 16);
    uint64_t *bar_full = s_bar, *bar_empty = s_bar + 4, *bar_dfull = s_bar + 8, *bar_dempty = s_bar + 10, *bar_table = s_bar + 12;

    const int tid = threadIdx.x, lane = tid & 31;
    // warp index through a shuffle: ptxas then knows it is warp-uniform, keeps the role branches uniform (BRA.U) and
    // the load descriptors / loop state in uniform registers instead of re-materialising them (R2UR) at every load
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    constexpr unsigned table_bytes = 2u * K * MG_N * 4u;

    if (tid == MG_PRODUCERS) {
        for (int k = 0; k < MG_SLOTS; k++) { mbar_init(&bar_full[k], MG_PRODUCERS); mbar_init(&bar_empty[k], 1); }
        for (int k = 0; k < 2; k++) { mbar_init(&bar_dfull[k], 1); mbar_init(&bar_dempty[k], MG_PRODUCERS); }
        mbar_init(bar_table, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "n"(MG_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *s_tmem;

    if (warp < 8) {
        // ===== producers + epilogue =====
        const int quarter = warp & 3, half = warp >> 2;
        const int m = quarter * 32 + lane;                                       // pixel inside the tile == TMEM lane
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
        const size_t plane = (size_t)(unsigned)hw * sizeof(float);
        const int my_cols = min(MG_HALF_N, Q - half * MG_HALF_N);                // query planes this thread stores (may be <= 0)

        // Register ring of RING stage buffers: the loads of flat stage k + RING are issued right after stage k has been
        // consumed, into the same registers.  For K = 256 (8 stages per tile) RING = 6: ~88 loads per thread in flight,
        // three quarters of a tile ahead; the ring position of a tile's first stage cycles through 0, 2, 4, so the tile
        // loop is unrolled three times (static register indices).  (RING = 8 needs 183 registers; 288 threads are
        // allocated as 12 warps, i.e. 168 registers per thread at most -- it spilled freshly loaded values.)
        constexpr int RING = (STAGES == 8) ? 6 : STAGES;
        constexpr int LAG = STAGES - RING;                                      // stages of the SAME tile still to prefetch
        float buf[RING][16];
        const char *qn = nullptr;                                                // running load pointer (tile being prefetched)
        char *o_cur = nullptr, *o_prev = nullptr, *o_next = nullptr;             // first output plane of a tile (null: row past the end)
        unsigned u = 0, t = 0;                                                   // stage uses / tiles so far (only parities matter)
        bool have_next = false, have_prev = false;

        TileCursor pf(blockIdx.x, n_items, slices, tiles_per_image);
        // thread's load pointer and output pointer for the cursor's tile
        auto setup = [&](const TileCursor &c, char *&o) {
            const long long p = c.first_pixel() + m;
            const long long pc = p < hw ? p : (long long)hw - 1;                 // rows past the end read the last pixel
            qn = reinterpret_cast<const char *>(feat + (c.b * K + half * 16) * (long long)hw + pc);
            o = (p < hw) ? reinterpret_cast<char *>(out + (c.b * Q + half * MG_HALF_N) * (long long)hw + p) : nullptr;
        };
        auto load_stage = [&](float (&dst)[16]) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                dst[i] = ldg_stream_f1(reinterpret_cast<const float *>(qn));
                qn += plane;
                asm volatile("" : "+l"(qn));     // keep ONE running pointer (ptxas otherwise precomputes and spills dozens)
            }
            qn += 16 * plane;                                                    // the other half's channels
            asm volatile("" : "+l"(qn));
        };
        // chunk c of the epilogue of tile t-1 (accumulator (t-1) & 1): 8 columns -> 8 query planes
        auto epi_chunk = [&](int c) {
            const unsigned tp = t - 1, db = tp & 1;
            if (c == 0) {
                mbar_wait(&bar_dfull[db], (tp >> 1) & 1);
                tc5_fence_after();
            }
            uint32_t v[8];
            tc5_ld8(lane_base + db * 128 + half * MG_HALF_N + c * 8, v);
            tc5_wait_ld();
            if (o_prev) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (c * 8 + i < my_cols) *reinterpret_cast<float *>(o_prev) = __uint_as_float(v[i]);
                    o_prev += plane;
                    asm volatile("" : "+l"(o_prev));
                }
            }
            if (c == MG_CHUNKS - 1) {
                tc5_fence_before();
                mbar_arrive(&bar_dempty[db]);                                    // the accumulator may be overwritten
            }
        };
        // one tile whose first stage sits at ring position OFF
        auto tile_body = [&](auto off_c) {
            constexpr int OFF = decltype(off_c)::value;
#pragma unroll
            for (int s = 0; s < STAGES; s++, u++) {
                float (&cur)[16] = buf[(OFF + s) % RING];
                const unsigned slot = u & (MG_SLOTS - 1);
                if (u >= MG_SLOTS) mbar_wait(&bar_empty[slot], ((u >> 2) + 1) & 1);   // MMAs of use u-4 are done
                tc5_fence_after();
                const uint32_t a = lane_base + MG_COL_A + slot * 64 + half * 16;
#pragma unroll
                for (int c = 0; c < 2; c++) {                                    // 8 channels at a time: 16 temporaries
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
                // flat stage + RING: the rest of this tile first, then the next tile
                if (s < LAG) load_stage(cur);
                else {
                    if (s == LAG) {
                        have_next = pf.valid();
                        if (have_next) { setup(pf, o_next); pf.next(); }
                    }
                    if (have_next) load_stage(cur);
                }
                if (have_prev) {
#pragma unroll
                    for (int c = s * CPS; c < (s + 1) * CPS && c < MG_CHUNKS; c++) epi_chunk(c);
                }
            }
            t++;
            o_prev = o_cur;
            have_prev = true;
            o_cur = o_next;
        };

        bool have = pf.valid();
        if (have) {
            setup(pf, o_cur);
            pf.next();
#pragma unroll
            for (int s = 0; s < RING; s++) load_stage(buf[s]);
        }
        while (have) {
            tile_body(std::integral_constant<int, 0>{});
            have = have_next;
            if (STAGES % RING != 0) {                                            // K = 256: ring positions 0, 2, 4
                if (!have) break;
                tile_body(std::integral_constant<int, (STAGES) % RING>{});
                have = have_next;
                if (!have) break;
                tile_body(std::integral_constant<int, (2 * STAGES) % RING>{});
                have = have_next;
            }
        }
        if (have_prev) {
#pragma unroll
            for (int c = 0; c < MG_CHUNKS; c++) epi_chunk(c);
        }
    } else {
        // ===== MMA issuer: the whole warp waits (stays converged), one elected lane issues =====
        const uint32_t bhi = smem_u32(s_b), blo = bhi + (uint32_t)K * MG_N * 4;
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);               // warp-uniform (see m2f_tc5q.cuh)
        unsigned u = 0, t = 0, n_loaded = 0;
        long long cur_item = -1;
        for (TileCursor c(blockIdx.x, n_items, slices, tiles_per_image); c.valid(); c.next(), t++) {
            if (c.item != cur_item) {
                // the image's table: every MMA that read the previous one has completed (commits are in order)
                if (t > 0) mbar_wait_backoff(&bar_dfull[(t - 1) & 1], ((t - 1) >> 1) & 1, 32);
                if (elect_one_sync()) {
                    mbar_expect_tx(bar_table, table_bytes);
                    bulk_load_1d(s_b, table + c.b * 2 * K * MG_N, table_bytes, bar_table);
                }
                __syncwarp();
                mbar_wait_backoff(bar_table, n_loaded & 1, 32);
                n_loaded++;
                cur_item = c.item;
            }
            const unsigned db = t & 1;
            if (t >= 2) mbar_wait_backoff(&bar_dempty[db], ((t >> 1) + 1) & 1, 32);   // epilogue of tile t-2 has read this accumulator
            const uint32_t d = tmem_u + db * 128;
#pragma unroll 1
            for (int s = 0; s < STAGES; s++, u++) {
                const unsigned slot = u & (MG_SLOTS - 1);
                mbar_wait_backoff(&bar_full[slot], (u >> 2) & 1, 32);
                tc5_fence_after();
                if (elect_one_sync()) {
#pragma unroll
                    for (int kk = 0; kk < MG_STAGE_K / 8; kk++) {
                        const int ks = s * (MG_STAGE_K / 8) + kk;                // k-step of 8 channels = 2 core-matrix chunks
                        const uint64_t dh = tc5_smem_desc(bhi + ks * 2 * (MG_N * 16), MG_N * 16, 128);
                        const uint64_t dl = tc5_smem_desc(blo + ks * 2 * (MG_N * 16), MG_N * 16, 128);
                        const uint32_t ahi = tmem_u + MG_COL_A + slot * 64 + kk * 8, alo = ahi + 32;
                        tc5_mma_ts(d, alo, dh, MG_IDESC, (s | kk) > 0);
                        tc5_mma_ts(d, ahi, dl, MG_IDESC, 1);
                        tc5_mma_ts(d, ahi, dh, MG_IDESC, 1);
                    }
                    tc5_commit(&bar_empty[slot]);
                    if (s == STAGES - 1) tc5_commit(&bar_dfull[db]);
                }
                __syncwarp();
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 8) {
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
    const int tiles_per_image = (int)((hw + 127) / 128);
    const int sms = sm_count();
    // B <= SMs: S = SMs / B slices per image, one work item per CTA; otherwise whole images, round-robin
    int slices = (B <= sms) ? std::min(tiles_per_image, sms / (int)B) : 1;
    if (slices < 1) slices = 1;
    const long long n_items = (long long)B * slices;
    const int grid = (int)std::min<long long>(n_items, (long long)sms);
    const size_t smem = mask_gemm_smem(K);
#define MG_LAUNCH(S)                                                                                                  \
    case S:                                                                                                           \
        MSS_CHECK_CUDA(cudaFuncSetAttribute(mask_gemm_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        mask_gemm_kernel<S><<<grid, MG_THREADS, smem, st>>>(mask_features, (int)hw, Q, n_items, slices, tiles_per_image, \
                                                            table, mask_logits);                                      \
        break;
    switch (K / MG_STAGE_K) {
        MG_LAUNCH(1) MG_LAUNCH(2) MG_LAUNCH(3) MG_LAUNCH(4) MG_LAUNCH(5) MG_LAUNCH(6) MG_LAUNCH(7) MG_LAUNCH(8)
    }
#undef MG_LAUNCH
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}
