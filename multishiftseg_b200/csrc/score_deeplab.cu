// (a1, a2) DeepLabv3+ fused per-pixel scoring: one read of the NCHW fp32 logits, up to four score
// maps written, optional fused evaluator append (ignore-label masking + key build).
//
// Replaces DeepWV3Plus.energy_func (lib/network/deepv3/deepv3.py:251-253):
//     anomaly_score = -(1. * torch.logsumexp(logit, dim=1))
// torch's logsumexp is  m = amax(x); m' = isinf(m) ? 0 : m;  log(sum exp(x - m')) + m'.
//
// HBM-bound: 4*C bytes read + 4 bytes per selected map written per pixel (C=19, three maps: 88 B/px).
// Layout: thread <-> 4 horizontally adjacent pixels; class c of those pixels is one 128-bit load from
// plane c, so a warp reads 512 contiguous bytes per plane and has C independent loads in flight.
#include "eval_append.cuh"

namespace mss {

struct ScoreOut {
    float *energy, *maxlogit, *msp, *entropy;
};

// exp(d) for d <= 0 (d = x - max): ex2.approx on the pre-scaled argument.  Relative error is 2^-22 from
// ex2.approx plus |d| * 6e-8 from rounding d*log2(e); terms with large |d| are negligible in the sum, so the
// sum's relative error stays ~1e-7 -- well inside the 1e-5 bar -- at 2 instructions instead of ~9 per class
// (ncu, round 1: the accurate expf kept issue utilisation at 72 % and capped the kernel at 84 % of HBM peak).
__device__ __forceinline__ float exp_neg(float d) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(d * 1.4426950408889634f));
    return e;
}

template <bool ENTROPY>
__device__ __forceinline__ void finish_pixel(float m, float s, float t, float &energy, float &maxlogit,
                                             float &msp, float &entropy, float msafe) {
    float ls = logf(s);
    energy = -(ls + msafe);
    maxlogit = -m;
    msp = 1.0f - 1.0f / s;
    if (ENTROPY) entropy = ls - t / s;
}

// ---- C known at compile time, 4 pixels per thread, 128-bit loads --------------------------------
template <int C, bool ENTROPY, bool EMIT>
__global__ void __launch_bounds__(256)
deeplab_score_vec4_kernel(const float *__restrict__ logits, long long HW, long long n_vec /* B*HW/4 */,
                          ScoreOut out, const void *__restrict__ labels, int label_dtype,
                          long long id_in, long long id_out, unsigned key_which, EvalDev ev) {
    const long long HW4 = HW >> 2;
    // grid-stride so EMIT's __syncthreads see uniform trip counts: round the loop bound up per CTA
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long iters = (n_vec + stride - 1) / stride;
    for (long long it = 0; it < iters; it++) {
        const long long v = first + it * stride;
        const bool active = v < n_vec;
        float sc[4] = {0.f, 0.f, 0.f, 0.f};
        long long pix = 0;
        unsigned mask = 0;
        AppendTicket ticket;
        if (EMIT) {
            // labels first: the CTA's range in the evaluator is reserved (one atomic) while the logits are scored
            if (active) {
                const long long b = v / HW4, p4 = v - b * HW4;
                mask = classify4(labels, label_dtype, b * HW + (p4 << 2), 4, id_in, id_out, true);
            }
            ticket = block_reserve4<256>(mask, ev);
        }
        if (active) {
            const long long b = v / HW4, p4 = v - b * HW4;
            const float *base = logits + (b * C) * HW + (p4 << 2);
            pix = b * HW + (p4 << 2);
            float4 x[C];
#pragma unroll
            for (int c = 0; c < C; c++) x[c] = ldg_stream_f4(base + (long long)c * HW);

            float4 m = x[0];
#pragma unroll
            for (int c = 1; c < C; c++) {
                m.x = fmaxf(m.x, x[c].x); m.y = fmaxf(m.y, x[c].y);
                m.z = fmaxf(m.z, x[c].z); m.w = fmaxf(m.w, x[c].w);
            }
            float4 ms;
            ms.x = isinf(m.x) ? 0.f : m.x; ms.y = isinf(m.y) ? 0.f : m.y;
            ms.z = isinf(m.z) ? 0.f : m.z; ms.w = isinf(m.w) ? 0.f : m.w;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f), t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < C; c++) {
                float dx = x[c].x - ms.x, dy = x[c].y - ms.y, dz = x[c].z - ms.z, dw = x[c].w - ms.w;
                float ex = exp_neg(dx), ey = exp_neg(dy), ez = exp_neg(dz), ew = exp_neg(dw);
                s.x += ex; s.y += ey; s.z += ez; s.w += ew;
                if (ENTROPY) {
                    // 0 * -inf guard: a -inf logit contributes p log p = 0
                    t.x += (ex == 0.f) ? 0.f : dx * ex; t.y += (ey == 0.f) ? 0.f : dy * ey;
                    t.z += (ez == 0.f) ? 0.f : dz * ez; t.w += (ew == 0.f) ? 0.f : dw * ew;
                }
            }
            float4 en, ml, mp, et;
            finish_pixel<ENTROPY>(m.x, s.x, t.x, en.x, ml.x, mp.x, et.x, ms.x);
            finish_pixel<ENTROPY>(m.y, s.y, t.y, en.y, ml.y, mp.y, et.y, ms.y);
            finish_pixel<ENTROPY>(m.z, s.z, t.z, en.z, ml.z, mp.z, et.z, ms.z);
            finish_pixel<ENTROPY>(m.w, s.w, t.w, en.w, ml.w, mp.w, et.w, ms.w);
            if (out.energy) stg_stream_f4(out.energy + pix, en);
            if (out.maxlogit) stg_stream_f4(out.maxlogit + pix, ml);
            if (out.msp) stg_stream_f4(out.msp + pix, mp);
            if (ENTROPY && out.entropy) stg_stream_f4(out.entropy + pix, et);
            if (EMIT) {
                float4 k = key_which == MSS_SCORE_ENERGY ? en
                         : key_which == MSS_SCORE_MAXLOGIT ? ml
                         : key_which == MSS_SCORE_MSP ? mp : et;
                sc[0] = k.x; sc[1] = k.y; sc[2] = k.z; sc[3] = k.w;
            }
        }
        if (EMIT) block_commit4<256>(sc, mask, ticket, ev);
    }
}

// ---- any C, any HW: one pixel per thread, online (max, sum, weighted sum) ------------------------
template <bool EMIT>
__global__ void __launch_bounds__(256)
deeplab_score_generic_kernel(const float *__restrict__ logits, int C, long long HW, long long n_pix,
                             ScoreOut out, const void *__restrict__ labels, int label_dtype,
                             long long id_in, long long id_out, unsigned key_which, EvalDev ev) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long iters = (n_pix + stride - 1) / stride;
    for (long long it = 0; it < iters; it++) {
        const long long i = first + it * stride;
        const bool active = i < n_pix;
        float sc[4] = {0.f, 0.f, 0.f, 0.f};
        if (active) {
            const long long b = i / HW, p = i - b * HW;
            const float *base = logits + (b * C) * HW + p;
            // pass 1: max (exactly torch's amax), pass 2: sums -- the second read hits L2 for the
            // plane strides used in practice; this path is the fallback, not the bench path
            float m = -INFINITY;
            for (int c = 0; c < C; c++) m = fmaxf(m, ldg_stream_f1(base + (long long)c * HW));
            float msafe = isinf(m) ? 0.f : m, s = 0.f, t = 0.f;
            for (int c = 0; c < C; c++) {
                float d = __ldg(base + (long long)c * HW) - msafe;
                float e = expf(d);
                s += e;
                t += (e == 0.f) ? 0.f : d * e;
            }
            float en, ml, mp, et;
            finish_pixel<true>(m, s, t, en, ml, mp, et, msafe);
            if (out.energy) out.energy[i] = en;
            if (out.maxlogit) out.maxlogit[i] = ml;
            if (out.msp) out.msp[i] = mp;
            if (out.entropy) out.entropy[i] = et;
            sc[0] = key_which == MSS_SCORE_ENERGY ? en
                  : key_which == MSS_SCORE_MAXLOGIT ? ml
                  : key_which == MSS_SCORE_MSP ? mp : et;
        }
        if (EMIT) {
            unsigned mask = active ? classify4(labels, label_dtype, i, 1, id_in, id_out, false) : 0u;
            block_append4<256>(sc, mask, ev);
        }
    }
}

// scores + labels -> evaluator (no logits): mss_eval_append
// (Round-2 experiment, not kept: 16 pixels per thread and reservation -- one pair of atomics and two barriers per 4096
// pixels instead of per 1024 -- was SLOWER, 678 vs 459 us for 134 M pixels: a thread then writes 16 consecutive keys, so
// one store instruction of a warp touches 32 sectors 64 bytes apart instead of 8 adjacent ones.)
__global__ void __launch_bounds__(256)
eval_append_kernel(const float *__restrict__ scores, const void *__restrict__ labels, int label_dtype,
                   long long n, long long id_in, long long id_out, EvalDev ev, int aligned) {
    const long long n4 = (n + 3) >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long iters = (n4 + stride - 1) / stride;
    for (long long it = 0; it < iters; it++) {
        const long long v = first + it * stride;
        float sc[4] = {0.f, 0.f, 0.f, 0.f};
        unsigned mask = 0;
        if (v < n4) {
            const long long i = v << 2;
            const int cnt = (n - i >= 4) ? 4 : (int)(n - i);
            if (aligned && cnt == 4) {
                float4 x = ldg_stream_f4(scores + i);
                sc[0] = x.x; sc[1] = x.y; sc[2] = x.z; sc[3] = x.w;
            } else {
                for (int j = 0; j < cnt; j++) sc[j] = scores[i + j];
            }
            mask = classify4(labels, label_dtype, i, cnt, id_in, id_out, aligned != 0);
        }
        block_append4<256>(sc, mask, ev);
    }
}

// The same for 16-byte-aligned inputs, built around what the round-2 profile of the kernel above showed at 134 M pixels:
// 2.5 TB/s, the SAME time for 1-byte and 8-byte labels, 2.1 warp instructions per pixel at 41 % issue utilisation and one
// store sector request per key -- instruction- and request-bound, not HBM-bound.  So:
//   * one reservation (atomic pair) and three barriers per 4096 pixels: a CTA takes four 1024-pixel sub-tiles at once
//     (16 pixels of loads in flight per thread) and ranks all four in two packed warp scans (4 x 8-bit counts per stream);
//   * labels are classified 4 at a time as byte flags (bit 8j+7 of a word = pixel j), exact zero-byte test for uint8;
//   * branch-free key / finiteness / compaction arithmetic (a NaN-or-Inf accumulator decided once per tile);
//   * the tile's keys are compacted in shared memory in the evaluator's own two-ended layout and copied out with
//     consecutive lanes storing consecutive keys (4-5 sectors per warp store instead of 16-32).
// (The 16-consecutive-pixels-per-thread form of the first idea alone was slower, 678 vs 459 us.)
constexpr int AW_SUB = 4, AW_TILE = 1024 * AW_SUB;      // sub-tiles of 1024 pixels per reservation

template <int LT> struct RawLabels;
template <> struct RawLabels<MSS_LABEL_U8> { unsigned w; };
template <> struct RawLabels<MSS_LABEL_I32> { int4 v; };
template <> struct RawLabels<MSS_LABEL_I64> { longlong2 a, b; };

__device__ __forceinline__ void load_raw(RawLabels<MSS_LABEL_U8> &r, const void *labels, long long i) {
    r.w = __ldcs(reinterpret_cast<const unsigned *>((const uint8_t *)labels + i));
}
__device__ __forceinline__ void load_raw(RawLabels<MSS_LABEL_I32> &r, const void *labels, long long i) {
    r.v = __ldcs(reinterpret_cast<const int4 *>((const int32_t *)labels + i));
}
__device__ __forceinline__ void load_raw(RawLabels<MSS_LABEL_I64> &r, const void *labels, long long i) {
    r.a = __ldcs(reinterpret_cast<const longlong2 *>((const long long *)labels + i));
    r.b = __ldcs(reinterpret_cast<const longlong2 *>((const long long *)labels + i + 2));
}

// label ids prepared once per thread.  uint8 labels: the id replicated into the 4 bytes of a word, or "never matches"
struct LabelIds {
    long long in, out;
    unsigned in4, out4;
    bool has_in8, has_out8;
};
__device__ __forceinline__ LabelIds make_ids(long long id_in, long long id_out) {
    LabelIds d;
    d.in = id_in; d.out = id_out;
    d.has_in8 = id_in >= 0 && id_in <= 255;
    d.has_out8 = id_out >= 0 && id_out <= 255;
    d.in4 = (unsigned)(id_in & 255) * 0x01010101u;
    d.out4 = (unsigned)(id_out & 255) * 0x01010101u;
    return d;
}
// 0x80 in every byte of x that is zero (exact, no borrow between bytes)
__device__ __forceinline__ unsigned zero_bytes(unsigned x) {
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}
// -> byte flags of 4 pixels: fin / fout have bit 8j+7 set when pixel j is in-distribution / OOD
__device__ __forceinline__ void flags_raw(const RawLabels<MSS_LABEL_U8> &r, const LabelIds &d, unsigned &fin, unsigned &fout) {
    fin = d.has_in8 ? zero_bytes(r.w ^ d.in4) : 0u;
    fout = d.has_out8 ? zero_bytes(r.w ^ d.out4) : 0u;
}
__device__ __forceinline__ void flags_of(long long l0, long long l1, long long l2, long long l3, const LabelIds &d,
                                         unsigned &fin, unsigned &fout) {
    const long long l[4] = {l0, l1, l2, l3};
    fin = fout = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (l[j] == d.in) fin |= 0x80u << (8 * j);
        if (l[j] == d.out) fout |= 0x80u << (8 * j);
    }
}
__device__ __forceinline__ void flags_raw(const RawLabels<MSS_LABEL_I32> &r, const LabelIds &d, unsigned &fin, unsigned &fout) {
    flags_of(r.v.x, r.v.y, r.v.z, r.v.w, d, fin, fout);
}
__device__ __forceinline__ void flags_raw(const RawLabels<MSS_LABEL_I64> &r, const LabelIds &d, unsigned &fin, unsigned &fout) {
    flags_of(r.a.x, r.a.y, r.b.x, r.b.y, d, fin, fout);
}

template <int LT>
__global__ void __launch_bounds__(256)
eval_append_wide_kernel(const float *__restrict__ scores, const void *__restrict__ labels, long long n, long long id_in,
                        long long id_out, EvalDev ev) {
    __shared__ unsigned s_cnt[8][2];                       // [warp][stream]: 4 x 8-bit counts, one byte per sub-tile
    __shared__ unsigned short s_off[8][AW_SUB][2];         // offset of (warp, sub-tile) inside the CTA's compacted tile
    __shared__ unsigned long long s_base[2];
    __shared__ unsigned s_tot[2];
    __shared__ uint32_t s_keys[AW_TILE + 256];             // + one dump slot per thread
    unsigned long long res_n = 0, res_p = 0;               // thread 0: reservation results
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long full = n >> 2, n4 = (n + 3) >> 2;      // whole groups of 4 pixels / groups incl. the ragged one
    const long long tiles = (n4 + 256 * AW_SUB - 1) / (256 * AW_SUB);
    const LabelIds ids = make_ids(id_in, id_out);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long g0 = tile * (256 * AW_SUB) + tid;
        const bool whole = (tile + 1) * (256 * AW_SUB) <= full;       // CTA-uniform: every group of the tile is complete
        float4 x[AW_SUB];
        unsigned valid[AW_SUB], posf[AW_SUB];              // byte flags (bit 8j+7) of the 4 pixels of each sub-tile
        if (whole) {
            RawLabels<LT> r[AW_SUB];
#pragma unroll
            for (int k = 0; k < AW_SUB; k++) {
                x[k] = ldg_stream_f4(scores + ((g0 + k * 256) << 2));
                load_raw(r[k], labels, (g0 + k * 256) << 2);
            }
#pragma unroll
            for (int k = 0; k < AW_SUB; k++) {
                unsigned fin;
                flags_raw(r[k], ids, fin, posf[k]);
                valid[k] = fin | posf[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < AW_SUB; k++) {             // the last tile: bounds-checked scalar loads
                const long long i = (g0 + k * 256) << 2;
                float sc[4];
                unsigned fin = 0, fout = 0;
                for (int j = 0; j < 4; j++) {
                    sc[j] = 0.f;
                    if (i + j < n) {
                        sc[j] = scores[i + j];
                        const long long l = load_label(labels, LT, i + j);
                        if (l == id_in) fin |= 0x80u << (8 * j);
                        if (l == id_out) fout |= 0x80u << (8 * j);
                    }
                }
                x[k] = make_float4(sc[0], sc[1], sc[2], sc[3]);
                posf[k] = fout;
                valid[k] = fin | fout;
            }
        }
        unsigned cn = 0, cp = 0;
#pragma unroll
        for (int k = 0; k < AW_SUB; k++) {
            const unsigned pc = __popc(posf[k]);
            cn |= (__popc(valid[k]) - pc) << (8 * k);
            cp |= pc << (8 * k);
        }
        unsigned in = cn, ip = cp;                         // inclusive warp scans, byte-wise (a warp has <= 128 per byte)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, in, d), w = __shfl_up_sync(0xffffffffu, ip, d);
            if (lane >= d) { in += u; ip += w; }
        }
        if (lane == 31) { s_cnt[warp][0] = in; s_cnt[warp][1] = ip; }
        __syncthreads();
        if (warp == 0) {
            const unsigned k = lane & 3, st = (lane >> 2) & 1; // lanes 0..7 own one (sub-tile, stream) each; the rest mirror
            unsigned run = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) run += (s_cnt[w][st] >> (8 * k)) & 255u;
            unsigned p = run;                              // inclusive prefix over the 4 sub-tiles of this stream
            unsigned t = __shfl_up_sync(0xffffffffu, p, 1);
            if (k >= 1) p += t;
            t = __shfl_up_sync(0xffffffffu, p, 2);
            if (k >= 2) p += t;
            if (lane < 8) {
                unsigned o = p - run;
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    s_off[w][k][st] = (unsigned short)o;
                    o += (s_cnt[w][st] >> (8 * k)) & 255u;
                }
            }
            const unsigned tn = __shfl_sync(0xffffffffu, p, 3), tp = __shfl_sync(0xffffffffu, p, 7);
            if (lane == 0) {
                s_tot[0] = tn;
                s_tot[1] = tp;
                if (tn) res_n = atomicAdd(&ev.state->n_neg, (unsigned long long)tn);   // consumed after the staging stores
                if (tp) res_p = atomicAdd(&ev.state->n_pos, (unsigned long long)tp);
            }
        }
        __syncthreads();
        // compact the tile's keys in shared memory in the evaluator's own layout: negatives up from 0, positives down from
        // the end
        unsigned amax = 0;                                 // max |score| bits over ALL 16 pixels (valid or not)
        {
            const unsigned exn = in - cn, exq = ip - cp;
            const unsigned dump = AW_TILE + tid;           // invalid pixels store to a private slot: no predicated stores
#pragma unroll
            for (int k = 0; k < AW_SUB; k++) {
                unsigned on = s_off[warp][k][0] + ((exn >> (8 * k)) & 255u);
                unsigned op = AW_TILE - 1 - (s_off[warp][k][1] + ((exq >> (8 * k)) & 255u));
                const float sc[4] = {x[k].x, x[k].y, x[k].z, x[k].w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const unsigned u = __float_as_uint(sc[j] + 0.0f);              // -0.0 -> +0.0: one key for both zeros
                    const unsigned key = u ^ (~(unsigned)((int)u >> 31) & 0x7FFFFFFFu);   // == score_key_desc
                    const unsigned v = (valid[k] >> (8 * j + 7)) & 1u, q = (posf[k] >> (8 * j + 7)) & 1u;
                    amax = max(amax, u & 0x7FFFFFFFu);
                    s_keys[v ? (q ? op : on) : dump] = key;
                    on += v - q;
                    op -= q;
                }
            }
        }
        if (amax >= 0x7F800000u) {                         // rare: some score is NaN / Inf -- sklearn raises if a VALID one is
            unsigned bad = 0;
#pragma unroll
            for (int k = 0; k < AW_SUB; k++) {
                const float sc[4] = {x[k].x, x[k].y, x[k].z, x[k].w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const unsigned a = __float_as_uint(sc[j]) & 0x7FFFFFFFu;
                    if (((valid[k] >> (8 * j + 7)) & 1u) && a >= 0x7F800000u) bad |= a > 0x7F800000u ? 1u : 2u;
                }
            }
            if (bad & 1u) atomicOr(&ev.state->nan_flag, 1u);
            if (bad & 2u) atomicOr(&ev.state->inf_flag, 1u);
        }
        if (tid == 0) {
            unsigned long long bn = res_n, bp = res_p;
            const unsigned tn = s_tot[0], tp = s_tot[1];
            // each stream is kept inside the buffer here; the two streams meeting is detected when the state is read
            if ((tn && bn + tn > (unsigned long long)ev.capacity) || (tp && bp + tp > (unsigned long long)ev.capacity)) {
                atomicAdd(&ev.state->overflow, (unsigned long long)(tn + tp));
                bn = bp = ~0ull;
            }
            s_base[0] = bn;
            s_base[1] = bp;
        }
        __syncthreads();
        // copy out: consecutive lanes store consecutive keys
        const unsigned long long bn = s_base[0], bp = s_base[1];
        if (bn != ~0ull) {
            const unsigned tn = s_tot[0], tp = s_tot[1];
            uint32_t *dn = ev.keys + bn, *dp = ev.keys + ((unsigned long long)ev.capacity - bp - tp);
            const uint32_t *sp = s_keys + (AW_TILE - tp);
#pragma unroll 4
            for (unsigned i = tid; i < tn; i += 256) dn[i] = s_keys[i];
#pragma unroll 4
            for (unsigned i = tid; i < tp; i += 256) dp[i] = sp[i];
        }
        // no barrier here: s_cnt is rewritten only by threads past this tile's second barrier (warp 0 has read it by then);
        // s_off / s_tot by warp 0 past the NEXT tile's first barrier and s_keys / s_base past its second and third -- every
        // thread has finished this tile's copy by then
    }
}

static EvalDev to_dev(const mss_eval_buffers *ev) {
    EvalDev d{nullptr, nullptr, 0};
    if (ev) { d.keys = ev->keys; d.state = (EvalState *)ev->state; d.capacity = ev->capacity; }
    return d;
}

static bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace mss

using namespace mss;

extern "C" int mss_eval_reset(const mss_eval_buffers *ev, void *stream) {
    MSS_REQUIRE(ev && ev->state, "mss_eval_reset: null evaluator");
    MSS_CHECK_CUDA(cudaMemsetAsync(ev->state, 0, MSS_EVAL_STATE_BYTES, (cudaStream_t)stream));
    return MSS_OK;
}

extern "C" int mss_eval_state_host(const mss_eval_buffers *ev, int64_t out_host[4], void *stream) {
    MSS_REQUIRE(ev && ev->state && out_host, "mss_eval_state_host: null argument");
    EvalState h;
    MSS_CHECK_CUDA(cudaMemcpyAsync(&h, ev->state, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MSS_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    out_host[0] = (int64_t)(h.n_neg + h.n_pos); out_host[1] = (int64_t)h.n_pos;
    out_host[2] = h.nan_flag; out_host[3] = h.inf_flag;
    if (h.overflow || (int64_t)(h.n_neg + h.n_pos) > ev->capacity) {       // the two streams met (or a CTA was dropped)
        set_error("evaluator capacity %lld exceeded (%llu valid pixels appended)", (long long)ev->capacity,
                  (unsigned long long)(h.n_neg + h.n_pos + h.overflow));
        return MSS_ERR_WORKSPACE;
    }
    return MSS_OK;
}

extern "C" int mss_eval_append(const float *scores, const void *labels, int label_dtype, int64_t n,
                               int64_t id_in, int64_t id_out, const mss_eval_buffers *ev, void *stream) {
    MSS_REQUIRE(ev && ev->keys && ev->state, "mss_eval_append: null evaluator buffers");
    MSS_REQUIRE(n >= 0, "mss_eval_append: n < 0");
    MSS_REQUIRE(label_dtype == MSS_LABEL_U8 || label_dtype == MSS_LABEL_I32 || label_dtype == MSS_LABEL_I64,
                "mss_eval_append: bad label_dtype %d", label_dtype);
    if (n == 0) return MSS_OK;
    MSS_REQUIRE(scores && labels, "mss_eval_append: null input");
    const int aligned = aligned16(scores) && aligned16(labels);
    long long n4 = (n + 3) / 4;
    int grid = (int)std::min<long long>((n4 + 255) / 256, (long long)sm_count() * 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (aligned) {
        // persistent grid of exactly the resident CTAs: every CTA runs the same number of tiles (+-1), no partial last wave
        static int per_sm[3] = {0, 0, 0};
        const int k = label_dtype == MSS_LABEL_U8 ? 0 : label_dtype == MSS_LABEL_I32 ? 1 : 2;
        if (!per_sm[k]) {
            int b = 0;
            const void *fn = k == 0 ? (const void *)eval_append_wide_kernel<MSS_LABEL_U8>
                           : k == 1 ? (const void *)eval_append_wide_kernel<MSS_LABEL_I32>
                                    : (const void *)eval_append_wide_kernel<MSS_LABEL_I64>;
            MSS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, fn, 256, 0));
            per_sm[k] = std::max(b, 1);
        }
        grid = (int)std::min<long long>((n4 + 256 * AW_SUB - 1) / (256 * AW_SUB), (long long)sm_count() * per_sm[k]);
    }
    if (!aligned)
        eval_append_kernel<<<grid, 256, 0, st>>>(scores, labels, label_dtype, n, id_in, id_out, to_dev(ev), 0);
    else if (label_dtype == MSS_LABEL_U8)
        eval_append_wide_kernel<MSS_LABEL_U8><<<grid, 256, 0, st>>>(scores, labels, n, id_in, id_out, to_dev(ev));
    else if (label_dtype == MSS_LABEL_I32)
        eval_append_wide_kernel<MSS_LABEL_I32><<<grid, 256, 0, st>>>(scores, labels, n, id_in, id_out, to_dev(ev));
    else
        eval_append_wide_kernel<MSS_LABEL_I64><<<grid, 256, 0, st>>>(scores, labels, n, id_in, id_out, to_dev(ev));
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

extern "C" int mss_deeplab_score(const float *logits, int64_t B, int C, int64_t HW, unsigned which,
                                 float *energy, float *maxlogit, float *msp, float *entropy,
                                 const void *labels, int label_dtype, int64_t id_in, int64_t id_out,
                                 unsigned key_which, const mss_eval_buffers *ev, void *stream) {
    MSS_REQUIRE(logits && B >= 0 && C >= 1 && HW >= 0, "mss_deeplab_score: bad shape B=%lld C=%d HW=%lld",
                (long long)B, C, (long long)HW);
    MSS_REQUIRE((which & ~15u) == 0 && which != 0, "mss_deeplab_score: bad `which` mask 0x%x", which);
    ScoreOut out{(which & MSS_SCORE_ENERGY) ? energy : nullptr, (which & MSS_SCORE_MAXLOGIT) ? maxlogit : nullptr,
                 (which & MSS_SCORE_MSP) ? msp : nullptr, (which & MSS_SCORE_ENTROPY) ? entropy : nullptr};
    MSS_REQUIRE(!(which & MSS_SCORE_ENERGY) || energy, "mss_deeplab_score: energy selected but NULL");
    MSS_REQUIRE(!(which & MSS_SCORE_MAXLOGIT) || maxlogit, "mss_deeplab_score: maxlogit selected but NULL");
    MSS_REQUIRE(!(which & MSS_SCORE_MSP) || msp, "mss_deeplab_score: msp selected but NULL");
    MSS_REQUIRE(!(which & MSS_SCORE_ENTROPY) || entropy, "mss_deeplab_score: entropy selected but NULL");
    const bool emit = ev != nullptr;
    if (emit) {
        MSS_REQUIRE(labels, "mss_deeplab_score: evaluator given without labels");
        MSS_REQUIRE(ev->keys && ev->state, "mss_deeplab_score: null evaluator buffers");
        MSS_REQUIRE(key_which && (key_which & (key_which - 1)) == 0 && (key_which & which),
                    "mss_deeplab_score: key_which must be exactly one of the selected scores");
        MSS_REQUIRE(label_dtype == MSS_LABEL_U8 || label_dtype == MSS_LABEL_I32 || label_dtype == MSS_LABEL_I64,
                    "mss_deeplab_score: bad label_dtype %d", label_dtype);
    }
    const long long n_pix = (long long)B * HW;
    if (n_pix == 0) return MSS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    EvalDev dev = to_dev(ev);
    const bool need_ent = (which & MSS_SCORE_ENTROPY) != 0;
    bool vec_ok = (C == 19) && (HW % 4 == 0) && aligned16(logits) && (!out.energy || aligned16(out.energy)) &&
                  (!out.maxlogit || aligned16(out.maxlogit)) && (!out.msp || aligned16(out.msp)) &&
                  (!out.entropy || aligned16(out.entropy)) && (!emit || aligned16(labels));
    if (vec_ok) {
        const long long n_vec = n_pix / 4;
        // plain one-shot grid unless the evaluator append needs uniform trip counts (grid-stride)
        long long blocks = (n_vec + 255) / 256;
        int grid = (int)std::min<long long>(blocks, emit ? (long long)sm_count() * 8 : blocks);
#define LAUNCH(ENT, EM)                                                                                  \
    deeplab_score_vec4_kernel<19, ENT, EM><<<grid, 256, 0, st>>>(logits, HW, n_vec, out, labels, label_dtype, \
                                                                 id_in, id_out, key_which, dev)
        if (need_ent) { if (emit) LAUNCH(true, true); else LAUNCH(true, false); }
        else          { if (emit) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
    } else {
        long long blocks = (n_pix + 255) / 256;
        int grid = (int)std::min<long long>(blocks, (long long)sm_count() * 32);
        if (emit)
            deeplab_score_generic_kernel<true><<<grid, 256, 0, st>>>(logits, C, HW, n_pix, out, labels, label_dtype,
                                                                     id_in, id_out, key_which, dev);
        else
            deeplab_score_generic_kernel<false><<<grid, 256, 0, st>>>(logits, C, HW, n_pix, out, labels, label_dtype,
                                                                      id_in, id_out, key_which, dev);
    }
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}
