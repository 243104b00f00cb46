// Device-side append of the order-preserving keys of VALID pixels to the streaming evaluator.
// Replaces the host accumulation of the reference tester loop (test_deeplab.py:94-101:
// .cpu().numpy() per batch + np.concatenate) and the label selection of metric.py:171-172.
//
// Two key-only streams in one buffer (round 2; round 1 stored (key, u8 label) pairs): label == id_in keys are
// appended upwards from keys[0], label == id_out keys downwards from keys[capacity - 1].  The label is the
// stream, so the sort, the multi-GPU exchange and the counting pass move 4 bytes per valid pixel and never a
// 1-byte store.  Order inside a stream is unspecified (the metric only depends on the two multisets).
#pragma once
#include "common.cuh"

namespace mss {

struct EvalDev {
    uint32_t *keys;
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

// CTA-aggregated append in two halves, so that the global atomics (one per stream and CTA, a ~0.5 us round trip to
// L2) overlap the scoring arithmetic instead of sitting between two barriers:
//   block_reserve4  labels only: classify, count, ONE atomicAdd per stream and CTA reserves the output ranges;
//                   thread 0 keeps the returned bases in registers and nobody waits for them yet
//   block_commit4   after the scores exist: publish the bases, write the keys
// Every thread of the CTA must call both (they contain __syncthreads).
// Counts travel packed: negatives in the low 16 bits, positives in the high 16 (a CTA holds <= 4 * BLOCK pixels).
struct AppendTicket {
    unsigned cnt, inc;             // this thread's (neg | pos << 16) counts, inclusive warp scan of them
    unsigned tot;                  // thread 0: counts of the CTA
    unsigned long long base_neg, base_pos;   // thread 0: reserved offsets (atomic results, possibly still in flight)
};

template <int BLOCK>
struct AppendSmem {
    unsigned warp_cnt[BLOCK / 32];
    unsigned long long cta_base[2];
};
template <int BLOCK>
__device__ __forceinline__ AppendSmem<BLOCK> &append_smem() {
    __shared__ AppendSmem<BLOCK> sm;
    return sm;
}

template <int BLOCK>
__device__ __forceinline__ AppendTicket block_reserve4(unsigned mask, const EvalDev &ev) {
    static_assert(BLOCK * 4 < 65536, "packed 16-bit counts");
    AppendSmem<BLOCK> &sm = append_smem<BLOCK>();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    AppendTicket t;
    const unsigned pos = __popc(mask >> 4);
    t.cnt = (__popc(mask & 15u) - pos) | (pos << 16);
    t.tot = 0;
    t.base_neg = t.base_pos = 0;
    unsigned inc = t.cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    t.inc = inc;
    if (lane == 31) sm.warp_cnt[warp] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
#pragma unroll
        for (int w = 0; w < BLOCK / 32; w++) {
            unsigned c = sm.warp_cnt[w];
            sm.warp_cnt[w] = tot;       // exclusive warp offsets (read by the other threads after commit's barrier)
            tot += c;
        }
        t.tot = tot;
        if (tot & 0xffffu) t.base_neg = atomicAdd(&ev.state->n_neg, (unsigned long long)(tot & 0xffffu));
        if (tot >> 16) t.base_pos = atomicAdd(&ev.state->n_pos, (unsigned long long)(tot >> 16));
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
        // each stream is kept inside the buffer here (memory safety); the two streams meeting in the middle
        // (n_neg + n_pos > capacity) is detected when the state is read (mss_eval_state_host)
        unsigned long long bn = t.base_neg, bp = t.base_pos;
        const unsigned tn = t.tot & 0xffffu, tp = t.tot >> 16;
        if ((tn && bn + tn > (unsigned long long)ev.capacity) || (tp && bp + tp > (unsigned long long)ev.capacity)) {
            atomicAdd(&ev.state->overflow, (unsigned long long)(tn + tp));
            bn = bp = ~0ull;        // drop: host reports MSS_ERR_WORKSPACE
        }
        sm.cta_base[0] = bn;
        sm.cta_base[1] = bp;
    }
    __syncthreads();
    const unsigned long long bn = sm.cta_base[0];
    if (bn != ~0ull && t.cnt) {
        const unsigned ex = sm.warp_cnt[warp] + (t.inc - t.cnt);       // packed exclusive offsets inside the CTA
        unsigned long long on = bn + (ex & 0xffffu);
        unsigned long long op = (unsigned long long)ev.capacity - 1 - (sm.cta_base[1] + (ex >> 16));
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if ((mask >> j) & 1u) {
                const uint32_t k = score_key_desc(s[j]);
                if ((mask >> (4 + j)) & 1u) ev.keys[op--] = k;
                else ev.keys[on++] = k;
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
