// Device-resident description of the two key streams a metric evaluation works on (negatives = in-distribution,
// positives = OOD), shared by the sort (radix_sort.cu) and the counting / tail stages (metrics.cu).
//
// The plan lives in DEVICE memory and is written by a one-thread kernel, either from host values or straight from the
// evaluator state.  Every later kernel reads the stream sizes from it, so the one-shot path
// (mss_ood_metrics: append -> sort -> counts -> ROC compaction) is enqueued without a single host round trip; grids are
// sized for an upper bound of the stream lengths and surplus CTAs exit.
#pragma once
#include "common.cuh"

namespace mss {

struct SortSeg {
    uint32_t *x;            // the caller's array: sorted in place
    uint32_t *y;            // alternate buffer (workspace)
    long long n;            // keys
    unsigned tile0, tiles;  // first tile (global numbering over both segments) and tile count of this segment
};

struct SortPlan {
    SortSeg seg[2];              // 0 = negatives (in-distribution), 1 = positives (OOD)
    unsigned total_tiles;
    unsigned char sel[2][4];     // pass p of segment s: 0 = x -> y, 1 = y -> x, 2 = skipped (all keys share the digit)
    unsigned char copy_back[2];  // the sorted segment ended in y: copy it to x
    unsigned pad;
};

size_t sort_ws_bytes(int64_t n_upper);

// write a plan for two already sorted (or to-be-merged) key arrays given from the host (no sort, y unused)
int plan_enqueue(SortPlan *plan_dev, const uint32_t *xa, int64_t na, const uint32_t *xb, int64_t nb, cudaStream_t st);

// Enqueue plan + digit histogram + four onesweep passes (+ copy back) for two independent key arrays.
//   ev != null : the streams are the evaluator's (sizes read from its device state; n_upper bounds n_neg + n_pos)
//   ev == null : (xa, na), (xb, nb) from the host (xb may be null / nb 0)
// *plan_dev receives the device address of the plan (inside ws); it stays valid until ws is reused.
int sort_enqueue(const mss_eval_buffers *ev, uint32_t *xa, int64_t na, uint32_t *xb, int64_t nb, int64_t n_upper,
                 void *ws, size_t ws_bytes, cudaStream_t st, const SortPlan **plan_dev);

}  // namespace mss
