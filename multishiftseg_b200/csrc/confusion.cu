// (8f-3) mIoU half of lib/utils/metric.py: the confusion histogram of hist_info (:10-18)
//     k = (gt >= 0) & (gt < n_cl);  labeled = sum(k);  correct = sum(pred[k] == gt[k])
//     hist = bincount(n_cl * gt[k] + pred[k], minlength=n_cl**2).reshape(n_cl, n_cl)
// accumulated on the device over a whole dataset (compute_metric's `hist += d['hist']`, :21-33).
// Integer work, bit-exact.  Two entry points: class-index maps (any of u8 / i32 / i64), and the fused
// form that takes the NCHW fp32 logits and does torch.argmax(dim=1) in registers (the logits are read
// once, the [B,H,W] int64 prediction map never exists).
//
// HBM-bound: pred + gt bytes per pixel (16 B/px for the reference's int64 maps), or 4*C + gt for the
// fused form.  Segmentation maps are spatially coherent, so a warp's 32 bins are usually 1-3 distinct
// values: each warp peels groups of equal bins (ballot + popc, one shared-memory atomic per group) and
// only falls back to per-lane atomics for what is left after 4 rounds.
#include "common.cuh"

namespace mss {

constexpr int CF_THREADS = 256;
constexpr int CF_MAX_CL = 32;

struct ConfDev {
    unsigned long long *hist;      // [n_cl * n_cl]  +=
    unsigned long long *lc;        // {labeled, correct, out_of_range}  +=
};

// add `valid ? 1 : 0` to s_hist[bin] for the 32 lanes of a warp
__device__ __forceinline__ void warp_hist_add(unsigned *s_hist, unsigned bin, bool valid) {
    unsigned remaining = __ballot_sync(0xffffffffu, valid);
    const unsigned lane = lane_id();
#pragma unroll 1
    for (int round = 0; round < 4 && remaining; round++) {
        const int leader = __ffs(remaining) - 1;
        const unsigned v = __shfl_sync(0xffffffffu, bin, leader);
        const unsigned m = __ballot_sync(0xffffffffu, valid && bin == v) & remaining;
        if ((int)lane == leader) atomicAdd(s_hist + v, (unsigned)__popc(m));
        remaining &= ~m;
    }
    if ((remaining >> lane) & 1u) atomicAdd(s_hist + bin, 1u);
}

// per-thread bookkeeping shared by both kernels: classify one (pred, gt) pair
struct PixelTally {
    unsigned labeled = 0, correct = 0, bad = 0;
    __device__ __forceinline__ bool classify(long long pred, long long gt, int n_cl, unsigned &bin) {
        bin = 0;
        if (gt < 0 || gt >= n_cl) return false;
        labeled++;
        correct += (pred == gt);
        const long long idx = (long long)n_cl * gt + pred;       // numpy: bincount(n_cl*gt + pred)
        if (idx < 0 || idx >= (long long)n_cl * n_cl) { bad++; return false; }   // numpy raises (negative / reshape)
        bin = (unsigned)idx;
        return true;
    }
};

__device__ __forceinline__ void flush_block(unsigned *s_hist, int bins, PixelTally t, ConfDev out) {
    // block totals of the three scalars
    __shared__ unsigned s_tot[3];
    if (threadIdx.x < 3) s_tot[threadIdx.x] = 0;
    __syncthreads();
    unsigned a = t.labeled, b = t.correct, c = t.bad;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, s);
        b += __shfl_xor_sync(0xffffffffu, b, s);
        c += __shfl_xor_sync(0xffffffffu, c, s);
    }
    if (lane_id() == 0) {
        if (a) atomicAdd(&s_tot[0], a);
        if (b) atomicAdd(&s_tot[1], b);
        if (c) atomicAdd(&s_tot[2], c);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += CF_THREADS)
        if (s_hist[i]) atomicAdd(out.hist + i, (unsigned long long)s_hist[i]);
    if (threadIdx.x < 3 && s_tot[threadIdx.x]) atomicAdd(out.lc + threadIdx.x, (unsigned long long)s_tot[threadIdx.x]);
}

// ---- class-index maps ---------------------------------------------------------------------------
// Each CTA owns chunks of <= 2^24 pixels between flushes, so the 32-bit shared counters cannot wrap.
__global__ void __launch_bounds__(CF_THREADS)
confusion_hist_kernel(const void *__restrict__ pred, int pred_dtype, const void *__restrict__ gt, int gt_dtype,
                      long long n, int n_cl, ConfDev out) {
    __shared__ unsigned s_hist[CF_MAX_CL * CF_MAX_CL];
    const int bins = n_cl * n_cl;
    for (int i = threadIdx.x; i < bins; i += CF_THREADS) s_hist[i] = 0;
    __syncthreads();
    PixelTally t;
    const long long stride = (long long)gridDim.x * CF_THREADS;
    const long long iters = (n + stride - 1) / stride;
    long long i = (long long)blockIdx.x * CF_THREADS + threadIdx.x;
    for (long long it = 0; it < iters; it++, i += stride) {
        unsigned bin = 0;
        bool valid = false;
        if (i < n) valid = t.classify(load_label(pred, pred_dtype, i), load_label(gt, gt_dtype, i), n_cl, bin);
        warp_hist_add(s_hist, bin, valid);
    }
    flush_block(s_hist, bins, t, out);
}

// ---- fused: NCHW fp32 logits -> argmax over C -> histogram ----------------------------------------
// torch.argmax: index of the FIRST maximal value; NaN counts as maximal (the first NaN wins).
__device__ __forceinline__ void argmax_step(float x, int c, float &best, int &arg) {
    if (x > best || (x != x && best == best)) { best = x; arg = c; }
}

template <int C>
__global__ void __launch_bounds__(CF_THREADS)
confusion_logits_vec4_kernel(const float *__restrict__ logits, long long HW, long long n_vec,
                             const void *__restrict__ gt, int gt_dtype, int n_cl, ConfDev out) {
    __shared__ unsigned s_hist[CF_MAX_CL * CF_MAX_CL];
    const int bins = n_cl * n_cl;
    for (int i = threadIdx.x; i < bins; i += CF_THREADS) s_hist[i] = 0;
    __syncthreads();
    PixelTally t;
    const long long HW4 = HW >> 2;
    const long long stride = (long long)gridDim.x * CF_THREADS;
    const long long iters = (n_vec + stride - 1) / stride;
    long long v = (long long)blockIdx.x * CF_THREADS + threadIdx.x;
    for (long long it = 0; it < iters; it++, v += stride) {
        unsigned bin[4] = {0, 0, 0, 0};
        bool valid[4] = {false, false, false, false};
        if (v < n_vec) {
            const long long b = v / HW4, p4 = v - b * HW4;
            const float *base = logits + (b * C) * HW + (p4 << 2);
            const long long pix = b * HW + (p4 << 2);
            float4 x[C];
#pragma unroll
            for (int c = 0; c < C; c++) x[c] = ldg_stream_f4(base + (long long)c * HW);
            long long g[4];
            if (gt_dtype == MSS_LABEL_U8) {
                const uchar4 q = *reinterpret_cast<const uchar4 *>((const uint8_t *)gt + pix);
                g[0] = q.x; g[1] = q.y; g[2] = q.z; g[3] = q.w;
            } else if (gt_dtype == MSS_LABEL_I32) {
                const int4 q = *reinterpret_cast<const int4 *>((const int32_t *)gt + pix);
                g[0] = q.x; g[1] = q.y; g[2] = q.z; g[3] = q.w;
            } else {
                const longlong2 q0 = *reinterpret_cast<const longlong2 *>((const long long *)gt + pix);
                const longlong2 q1 = *reinterpret_cast<const longlong2 *>((const long long *)gt + pix + 2);
                g[0] = q0.x; g[1] = q0.y; g[2] = q1.x; g[3] = q1.y;
            }
            float best[4] = {x[0].x, x[0].y, x[0].z, x[0].w};
            int arg[4] = {0, 0, 0, 0};
#pragma unroll
            for (int c = 1; c < C; c++) {
                argmax_step(x[c].x, c, best[0], arg[0]);
                argmax_step(x[c].y, c, best[1], arg[1]);
                argmax_step(x[c].z, c, best[2], arg[2]);
                argmax_step(x[c].w, c, best[3], arg[3]);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) valid[j] = t.classify(arg[j], g[j], n_cl, bin[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) warp_hist_add(s_hist, bin[j], valid[j]);
    }
    flush_block(s_hist, bins, t, out);
}

// any C / HW not a multiple of 4 / unaligned: one pixel per thread
__global__ void __launch_bounds__(CF_THREADS)
confusion_logits_generic_kernel(const float *__restrict__ logits, int C, long long HW, long long n_pix,
                                const void *__restrict__ gt, int gt_dtype, int n_cl, ConfDev out) {
    __shared__ unsigned s_hist[CF_MAX_CL * CF_MAX_CL];
    const int bins = n_cl * n_cl;
    for (int i = threadIdx.x; i < bins; i += CF_THREADS) s_hist[i] = 0;
    __syncthreads();
    PixelTally t;
    const long long stride = (long long)gridDim.x * CF_THREADS;
    const long long iters = (n_pix + stride - 1) / stride;
    long long i = (long long)blockIdx.x * CF_THREADS + threadIdx.x;
    for (long long it = 0; it < iters; it++, i += stride) {
        unsigned bin = 0;
        bool valid = false;
        if (i < n_pix) {
            const long long b = i / HW, p = i - b * HW;
            const float *base = logits + (b * C) * HW + p;
            float best = ldg_stream_f1(base);
            int arg = 0;
            for (int c = 1; c < C; c++) argmax_step(ldg_stream_f1(base + (long long)c * HW), c, best, arg);
            valid = t.classify(arg, load_label(gt, gt_dtype, i), n_cl, bin);
        }
        warp_hist_add(s_hist, bin, valid);
    }
    flush_block(s_hist, bins, t, out);
}

static int conf_grid(long long work_items) {
    // <= 2^24 items per CTA-lifetime keeps the 32-bit shared counters exact for any n below 2^24 * grid;
    // beyond that the host loop in the entry points splits the call
    const long long want = (work_items + CF_THREADS - 1) / CF_THREADS;
    return (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count() * 8));
}

static bool label_dtype_ok(int d) { return d == MSS_LABEL_U8 || d == MSS_LABEL_I32 || d == MSS_LABEL_I64; }

}  // namespace mss

using namespace mss;

extern "C" int mss_confusion_hist(const void *pred, int pred_dtype, const void *gt, int gt_dtype, int64_t n, int n_cl,
                                  int64_t *hist, int64_t *labeled_correct, void *stream) {
    MSS_REQUIRE(n >= 0 && n_cl >= 1 && n_cl <= CF_MAX_CL, "mss_confusion_hist: need n >= 0 and 1 <= n_cl <= 32");
    MSS_REQUIRE(hist && labeled_correct, "mss_confusion_hist: null accumulator");
    MSS_REQUIRE(label_dtype_ok(pred_dtype) && label_dtype_ok(gt_dtype), "mss_confusion_hist: dtype must be u8 / i32 / i64");
    ConfDev out{(unsigned long long *)hist, (unsigned long long *)labeled_correct};
    if (n == 0) return MSS_OK;
    MSS_REQUIRE(pred && gt, "mss_confusion_hist: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t chunk = (int64_t)1 << 32;            // 2^32 px over >= 256 CTAs <= 2^24 per CTA
    for (int64_t o = 0; o < n; o += chunk) {
        const int64_t m = std::min<int64_t>(chunk, n - o);
        confusion_hist_kernel<<<m >= chunk ? sm_count() * 8 : conf_grid(m), CF_THREADS, 0, st>>>(
            (const char *)pred + o * pred_dtype, pred_dtype, (const char *)gt + o * gt_dtype, gt_dtype, m, n_cl, out);
        MSS_CHECK_LAUNCH();
    }
    return MSS_OK;
}

extern "C" int mss_confusion_from_logits(const float *logits, int64_t B, int C, int64_t HW, const void *gt,
                                         int gt_dtype, int n_cl, int64_t *hist, int64_t *labeled_correct,
                                         void *stream) {
    MSS_REQUIRE(B >= 0 && C >= 1 && HW >= 0, "mss_confusion_from_logits: bad shape");
    MSS_REQUIRE(n_cl >= 1 && n_cl <= CF_MAX_CL, "mss_confusion_from_logits: need 1 <= n_cl <= 32");
    MSS_REQUIRE(hist && labeled_correct, "mss_confusion_from_logits: null accumulator");
    MSS_REQUIRE(label_dtype_ok(gt_dtype), "mss_confusion_from_logits: gt dtype must be u8 / i32 / i64");
    ConfDev out{(unsigned long long *)hist, (unsigned long long *)labeled_correct};
    const int64_t n = B * HW;
    if (n == 0) return MSS_OK;
    MSS_REQUIRE(logits && gt, "mss_confusion_from_logits: null pointer");
    MSS_REQUIRE(n < ((int64_t)1 << 32), "mss_confusion_from_logits: more than 2^32 pixels per call");
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = C == 19 && (HW & 3) == 0 && (((uintptr_t)logits) & 15) == 0 &&
                     (((uintptr_t)gt) & (gt_dtype == MSS_LABEL_I64 ? 15 : 4 * gt_dtype - 1)) == 0;
    if (vec) {
        confusion_logits_vec4_kernel<19><<<conf_grid(n / 4), CF_THREADS, 0, st>>>(logits, HW, n / 4, gt, gt_dtype, n_cl, out);
    } else {
        confusion_logits_generic_kernel<<<conf_grid(n), CF_THREADS, 0, st>>>(logits, C, HW, n, gt, gt_dtype, n_cl, out);
    }
    MSS_CHECK_LAUNCH();
    return MSS_OK;
}

// accumulators -> host: the ONE synchronisation of a streaming evaluation (the update calls above only enqueue)
extern "C" int mss_confusion_result(const int64_t *hist, const int64_t *labeled_correct, int n_cl, int64_t *hist_host,
                                    int64_t labeled_correct_host[2], void *stream) {
    MSS_REQUIRE(hist && labeled_correct && hist_host && labeled_correct_host && n_cl >= 1 && n_cl <= CF_MAX_CL,
                "mss_confusion_result: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long long lc[3] = {0, 0, 0};
    MSS_CHECK_CUDA(cudaMemcpyAsync(hist_host, hist, (size_t)n_cl * n_cl * 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(lc, labeled_correct, sizeof(lc), cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    labeled_correct_host[0] = lc[0];
    labeled_correct_host[1] = lc[1];
    if (lc[2]) {
        // numpy's bincount / reshape raise on the first batch that has such a pixel (metric.py:15-17)
        set_error("mss_confusion_result: %lld labeled pixel(s) with n_cl*gt + pred outside [0, n_cl^2) (numpy's bincount/reshape raises)",
                  lc[2]);
        return MSS_ERR_INVALID_ARG;
    }
    return MSS_OK;
}

// np.nanmean over a short float64 vector: NaNs count as 0 in the sum (numpy's pairwise order) and not in the divisor
static double np_nanmean(const double *a, int n) {
    double tmp[CF_MAX_CL];
    int cnt = 0;
    for (int i = 0; i < n; i++) {
        const bool is_nan = a[i] != a[i];
        tmp[i] = is_nan ? 0.0 : a[i];
        cnt += !is_nan;
    }
    double tot = 0.0;
    if (n > 0) mss_pairwise_sum_host(tmp, n, &tot);
    volatile double r = tot / (double)cnt;       // 0 / 0 -> nan ("Mean of empty slice" in numpy)
    return r;
}

// host-only: the closed-form arithmetic of compute_score (metric.py:42-49; per_class = 0) and compute_score_per_class
// (metric.py:51-64; per_class = 1) on the accumulated confusion matrix, IEEE float64 operation by operation as numpy
// evaluates it.  hist_host is the float64 accumulator of compute_metric (metric.py:22), row = gt, column = pred.
//   iu_host [n_cl]          per-class IoU
//   class_acc_host [n_cl]   per-class accuracy (per_class = 1 only; may be NULL otherwise)
//   out_host = {mean_IU, mean_IU_no_back (per_class = 0) or nan, mean_pixel_acc}
extern "C" int mss_confusion_scores(const double *hist_host, int n_cl, double correct, double labeled, int per_class,
                                    double *iu_host, double *class_acc_host, double out_host[3]) {
    MSS_REQUIRE(hist_host && iu_host && out_host && n_cl >= 1 && n_cl <= CF_MAX_CL, "mss_confusion_scores: bad arguments");
    MSS_REQUIRE(!per_class || class_acc_host, "mss_confusion_scores: per_class needs class_acc_host");
    for (int c = 0; c < n_cl; c++) {
        double row = 0.0, col = 0.0;             // integer-valued float64 sums: exact in any order
        for (int j = 0; j < n_cl; j++) { row += hist_host[c * n_cl + j]; col += hist_host[j * n_cl + c]; }
        const double inter = hist_host[c * n_cl + c];
        volatile double uni = row + col;
        uni = uni - inter;
        if (per_class) {
            volatile double q = inter / (uni > 1.0 ? (double)uni : 1.0);
            iu_host[c] = q;
            volatile double a = inter / (row > 1.0 ? row : 1.0);
            class_acc_host[c] = a;
        } else {
            volatile double q = inter / uni;     // 0 / 0 -> nan, skipped by nanmean
            iu_host[c] = q;
        }
    }
    out_host[0] = np_nanmean(iu_host, n_cl);
    out_host[1] = per_class ? nan("") : np_nanmean(iu_host + 1, n_cl - 1);
    volatile double acc = correct / labeled;
    out_host[2] = acc;
    return MSS_OK;
}
