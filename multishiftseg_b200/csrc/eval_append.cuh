// Device-side append of (key, label) pairs of VALID pixels to the streaming evaluator.
// Replaces the host accumulation of the reference tester loop (test_deeplab.py:94-101:
// .cpu().numpy() per batch + np.concatenate) and the label selection of metric.py:171-172.
// Order inside the evaluator buffer is unspecified (the metric only depends on the multiset).
#pragma once
#include "common.cuh"

namespace mss {

struct EvalDev {
    uint32_t *keys;
    uint8_t *labs;
    EvalState *state;
    long long capacity;
};

// Load 4 consecutive labels starting at pixel index i (i % 4 == 0 and base suitably aligned when
// `aligned`), classify each as in (0) / out (1) / ignored.  Returns bit masks: valid | pos << 4.
__device__ __forceinline__ unsigned classify4(const void *labels, int dtype, long long i, int nvalid_px,
                                              long long id_in, long long id_out, bool aligned) {
    long long l[4];
    if (aligned && nvalid_px == 4) {
        if (dtype == MSS_LABEL_U8) {
            uchar4 v = *reinterpret_cast<const uchar4 *>((const uint8_t *)labels + i);
            l[0] = v.x; l[1] = v.y; l[2] = v.z; l[3] = v.w;
        } else if (dtype == MSS_LABEL_I32) {
            int4 v = *reinterpret_cast<const int4 *>((const int32_t *)labels + i);
            l[0] = v.x; l[1] = v.y; l[2] = v.z; l[3] = v.w;
        } else {
            longlong2 a = *reinterpret_cast<const longlong2 *>((const long long *)labels + i);
            longlong2 b = *reinterpret_cast<const longlong2 *>((const long long *)labels + i + 2);
            l[0] = a.x; l[1] = a.y; l[2] = b.x; l[3] = b.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) l[j] = (j < nvalid_px) ? load_label(labels, dtype, i + j) : -1;
    }
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (j < nvalid_px) {
            bool pos = (l[j] == id_out);
            bool in = (l[j] == id_in);
            if (pos || in) m |= 1u << j;
            if (pos) m |= 16u << j;
        }
    }
    return m;
}

// CTA-aggregated append in two halves, so that the one global atomic per CTA (a ~0.5 us round trip to L2) overlaps the
// scoring arithmetic instead of sitting between two barriers: with ~100 registers per thread only two scoring CTAs fit on
// an SM, and the serial form (score, barrier, atomic, barrier, write) left the fused kernel at 81 % of the HBM roofline
// against 97 % for scoring alone.
//   block_reserve4  labels only: classify, count, ONE atomicAdd per CTA reserves the output range; thread 0 keeps the
//                   returned base in a register and nobody waits for it yet
//   block_commit4   after the scores exist: publish the base, write (key, label) pairs
// Every thread of the CTA must call both (they contain __syncthreads).
// (Measured, round 1, cfg-4 scoring+append: serial form 68.1 ms, this split 65.6 ms; additionally compacting each warp's
// pairs in shared memory so that consecutive lanes store consecutive elements: 66.3 ms -- not kept.)
struct AppendTicket {
    unsigned cnt, inc;             // this thread's valid pixels, inclusive warp scan of them
    unsigned tot;                  // thread 0: valid pixels of the CTA
    unsigned long long base;       // thread 0: reserved offset (atomic result, possibly still in flight)
};

template <int BLOCK>
struct AppendSmem {
    unsigned warp_cnt[BLOCK / 32];
    unsigned warp_pos[BLOCK / 32];
    unsigned long long cta_base;
};
template <int BLOCK>
__device__ __forceinline__ AppendSmem<BLOCK> &append_smem() {
    __shared__ AppendSmem<BLOCK> sm;
    return sm;
}

template <int BLOCK>
__device__ __forceinline__ AppendTicket block_reserve4(unsigned mask, const EvalDev &ev) {
    AppendSmem<BLOCK> &sm = append_smem<BLOCK>();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    AppendTicket t;
    t.cnt = __popc(mask & 15u);
    t.tot = 0;
    t.base = 0;
    const unsigned pos = __popc(mask >> 4);
    // warp inclusive scan of cnt, warp sum of pos
    unsigned inc = t.cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    t.inc = inc;
    const unsigned psum = __reduce_add_sync(0xffffffffu, pos);
    if (lane == 31) { sm.warp_cnt[warp] = inc; sm.warp_pos[warp] = psum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0, ptot = 0;
#pragma unroll
        for (int w = 0; w < BLOCK / 32; w++) {
            unsigned c = sm.warp_cnt[w];
            sm.warp_cnt[w] = tot;       // exclusive warp offsets (read by the other threads after commit's barrier)
            tot += c;
            ptot += sm.warp_pos[w];
        }
        t.tot = tot;
        if (tot) {
            t.base = atomicAdd(&ev.state->count, (unsigned long long)tot);
            if (ptot) atomicAdd(&ev.state->n_pos, (unsigned long long)ptot);
        }
    }
    return t;
}

// `mask` as returned by classify4; s[j] the score of pixel j.
template <int BLOCK>
__device__ __forceinline__ void block_commit4(const float s[4], unsigned mask, const AppendTicket &t, const EvalDev &ev) {
    AppendSmem<BLOCK> &sm = append_smem<BLOCK>();
    const unsigned warp = threadIdx.x >> 5;

    // non-finite valid scores: sklearn raises (assert_all_finite, _ranking.py:896-897)
    unsigned bad = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if ((mask >> j) & 1u) {
            unsigned u = __float_as_uint(s[j]) & 0x7FFFFFFFu;
            if (u > 0x7F800000u) bad |= 1u;
            else if (u == 0x7F800000u) bad |= 2u;
        }
    }
    if (bad & 1u) atomicOr(&ev.state->nan_flag, 1u);
    if (bad & 2u) atomicOr(&ev.state->inf_flag, 1u);

    if (threadIdx.x == 0) {
        unsigned long long base = t.base;
        if (t.tot && base + t.tot > (unsigned long long)ev.capacity) {
            atomicAdd(&ev.state->overflow, (unsigned long long)t.tot);
            base = ~0ull;        // drop: host reports MSS_ERR_WORKSPACE
        }
        sm.cta_base = base;
    }
    __syncthreads();
    const unsigned long long base = sm.cta_base;
    if (base != ~0ull && t.cnt) {
        unsigned long long o = base + sm.warp_cnt[warp] + (t.inc - t.cnt);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if ((mask >> j) & 1u) {
                ev.keys[o] = score_key_desc(s[j]);
                ev.labs[o] = (uint8_t)((mask >> (4 + j)) & 1u);
                o++;
            }
        }
    }
    __syncthreads();   // shared scratch is reused by the next call of a grid-stride loop
}

// both halves back to back (callers that have nothing to overlap)
template <int BLOCK>
__device__ __forceinline__ void block_append4(const float s[4], unsigned mask, const EvalDev &ev) {
    const AppendTicket t = block_reserve4<BLOCK>(mask, ev);
    block_commit4<BLOCK>(s, mask, t, ev);
}

}  // namespace mss
