// Shared device / host helpers for libmss_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/mss_b200.h"

namespace mss {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define MSS_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            mss::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                           __LINE__);                                                          \
            return MSS_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define MSS_CHECK_LAUNCH()                                                                     \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            mss::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),         \
                           __FILE__, __LINE__);                                                \
            return MSS_ERR_CUDA;                                                               \
        }                                                                                      \
        mss::count_launch();                                                                   \
    } while (0)

#define MSS_REQUIRE(cond, ...)                                                                 \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            mss::set_error(__VA_ARGS__);                                                       \
            return MSS_ERR_INVALID_ARG;                                                        \
        }                                                                                      \
    } while (0)

int sm_count();  // SMs of the current device (148 on B200), cached per device

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: a call site keeps one bit per device
// ordinal (a process may drive several GPUs, e.g. nn.DataParallel threads) instead of one process-wide flag.
inline bool first_use_on_device(std::atomic<unsigned long long> &seen) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return true;
    const unsigned long long bit = 1ull << (dev & 63);
    if (seen.load(std::memory_order_acquire) & bit) return false;
    seen.fetch_or(bit, std::memory_order_acq_rel);     // (two threads may both set the attribute once: harmless)
    return true;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over a caller-provided workspace
struct Carver {
    char *base;
    size_t off = 0, cap;
    Carver(void *p, size_t bytes) : base((char *)p), cap(bytes) {}
    template <typename T>
    T *take(size_t n) {
        off = align_up(off, 256);
        T *r = (T *)(base + off);
        off += n * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

// ---- evaluator state (device) -----------------------------------------------------------------
// Two key-only streams share ONE key buffer: in-distribution (negative) keys fill keys[0, n_neg) upwards, OOD
// (positive) keys fill keys[capacity - n_pos, capacity) downwards.  The 0/1 label is the stream a key lives in, so no
// label byte is ever stored, sorted or exchanged.
struct EvalState {
    unsigned long long n_neg;  // in-distribution pixels appended so far
    unsigned long long n_pos;  // OOD pixels appended so far
    unsigned int nan_flag;
    unsigned int inf_flag;
    unsigned long long overflow;  // appends dropped because capacity was exceeded
    unsigned long long pad[4];
};
static_assert(sizeof(EvalState) == MSS_EVAL_STATE_BYTES, "EvalState size is ABI");

#ifdef __CUDACC__
// ---- device helpers ---------------------------------------------------------------------------
// streaming 128-bit global load: read-only path, do not allocate in L1 (data is touched once)
__device__ __forceinline__ float4 ldg_stream_f4(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream_f1(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream_f4(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void stg_stream_f1(float *p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// float32 score -> uint32 key whose ascending order is the descending score order;
// -0.0 and +0.0 collapse to one key (np.diff(y_score) == 0 between them: one threshold).
__device__ __forceinline__ uint32_t score_key_desc(float f) {
    uint32_t u = __float_as_uint(f);
    if (u == 0x80000000u) u = 0u;
    uint32_t asc = (u >> 31) ? ~u : (u | 0x80000000u);
    return ~asc;
}

__device__ __forceinline__ long long load_label(const void *labels, int dtype, long long i) {
    if (dtype == MSS_LABEL_U8) return ((const uint8_t *)labels)[i];
    if (dtype == MSS_LABEL_I32) return ((const int32_t *)labels)[i];
    return ((const long long *)labels)[i];
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// same, for a warp that has nothing else to do (the MMA issuer): back off between polls so the spin does not
// take issue slots from the warps doing the arithmetic
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, unsigned parity, unsigned ns) {
    for (;;) {
        unsigned done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map): src/dst 16-byte aligned, bytes % 16 == 0;
// completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_one(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---- decoupled look-back status words (sort passes, counting passes) ------------------------------------------
// bits 63..62 = 0 empty / 1 tile aggregate / 2 inclusive prefix, low 62 bits = value
constexpr unsigned long long FLAG_AGG = 1ull << 62, FLAG_INC = 2ull << 62, VAL_MASK = (1ull << 62) - 1;
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// One value per tile (compaction offsets).  The exclusive prefix of tile t does not depend on tile t's own count, so
// the walk over the predecessors' status words is done by a DEDICATED warp that starts right after the ticket is
// drawn, while the other warps still load and process the tile (ncu, round 2, first form -- warp 0 walked after the
// CTA's scan with everyone else parked at a barrier: ~9 polls of ~0.5 us per tile, 45-55 % of all stall samples
// on that barrier).  Protocol per tile:
//   worker thread : tile_publish_aggregate(status, tile, tot)       as soon as the tile's count is known
//   walker warp   : ex = tile_walk(status, tile)                    all 32 lanes; 64 predecessors per round trip
//   after a CTA barrier, walker lane 0: tile_publish_inclusive(status, tile, ex + tot)
// Tiles must be numbered in start order (atomic ticket) so that a walk only ever waits on running tiles.
__device__ __forceinline__ void tile_publish_aggregate(unsigned long long *status, unsigned tile, unsigned long long tot) {
    if (tile > 0) st_status(status + tile, FLAG_AGG | tot);       // tile 0 goes straight to "inclusive"
}
__device__ __forceinline__ void tile_publish_inclusive(unsigned long long *status, unsigned tile, unsigned long long inc) {
    st_status(status + tile, FLAG_INC | inc);
}
__device__ __forceinline__ unsigned long long tile_walk(const unsigned long long *status, unsigned tile) {
    const unsigned lane = threadIdx.x & 31;
    if (tile == 0) return 0;
    unsigned long long prefix = 0;
    long long base = (long long)tile - 1;
    for (;;) {
        // near window: tiles base .. base-31 (lane order), far window: base-32 .. base-63
        const long long i0 = base - lane, i1 = base - 32 - lane;
        const unsigned long long s0 = i0 >= 0 ? ld_status(status + i0) : FLAG_INC;       // before tile 0: prefix 0
        const unsigned long long s1 = i1 >= 0 ? ld_status(status + i1) : FLAG_INC;
        const unsigned f0 = (unsigned)(s0 >> 62), f1 = (unsigned)(s1 >> 62);
        const unsigned inc0 = __ballot_sync(0xffffffffu, f0 == 2u), emp0 = __ballot_sync(0xffffffffu, f0 == 0u);
        const unsigned upto0 = inc0 ? ((2u << (__ffs(inc0) - 1)) - 1u) : 0xffffffffu;    // lanes up to the first inclusive one
        if (emp0 & upto0) { __nanosleep(64); continue; }                                 // a needed predecessor has not published yet
        unsigned long long v = ((upto0 >> lane) & 1u) ? (s0 & VAL_MASK) : 0ull;
        bool done = inc0 != 0;
        bool advance64 = false;
        if (!done) {
            const unsigned inc1 = __ballot_sync(0xffffffffu, f1 == 2u), emp1 = __ballot_sync(0xffffffffu, f1 == 0u);
            const unsigned upto1 = inc1 ? ((2u << (__ffs(inc1) - 1)) - 1u) : 0xffffffffu;
            if (!(emp1 & upto1)) {                                                       // far window usable as well
                v += ((upto1 >> lane) & 1u) ? (s1 & VAL_MASK) : 0ull;
                done = inc1 != 0;
                advance64 = true;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        prefix += v;
        if (done) break;
        base -= advance64 ? 64 : 32;
    }
    return prefix;
}
// CTA barrier over a subset of the warps (the workers): barrier resource `id` (1..15), `threads` a multiple of 32
__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
#endif  // __CUDACC__

}  // namespace mss
