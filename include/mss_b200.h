/* mss_b200 -- C ABI of the B200-native dense anomaly-scoring + exact OOD-evaluation path.
 *
 * The reference (gaozhitong/MultiShiftSeg) has no plugin / FFI layer: its boundary for this
 * path is a handful of Python functions.  Each entry point below replaces one of them (cited
 * as file:line relative to the reference root) and is what a ctypes / cffi / pybind stub on
 * the reference side would bind.
 *
 * Conventions
 *   - plain pointers and sizes only; every tensor pointer is a DEVICE pointer on the current
 *     CUDA device unless the parameter name ends in `_host`;
 *   - the caller owns every buffer, including workspaces; the library keeps no global mutable
 *     state besides a thread-local error string, and is re-entrant (nn.DataParallel calls the
 *     scoring functions from one thread per GPU);
 *   - every launch goes to the `stream` argument (a cudaStream_t passed as void*);
 *   - return value: 0 = MSS_OK, >0 = a non-error outcome the reference also has
 *     (MSS_EMPTY_CLASS <-> `eval_ood_measure` returning None), <0 = error; never throws/aborts.
 *     `mss_last_error()` returns a thread-local human-readable message for the last <0 code.
 *   - layouts are the reference's: NCHW contiguous fp32 logits, [B,Q,C+1] class logits,
 *     [B,Q,h,w] mask logits, row-major score / label maps.
 */
#ifndef MSS_B200_H
#define MSS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSS_ABI_VERSION 2

#if defined(__GNUC__)
#define MSS_API __attribute__((visibility("default")))
#else
#define MSS_API
#endif

/* return codes */
#define MSS_OK 0
#define MSS_EMPTY_CLASS 1        /* no ID or no OOD pixel: reference returns None (metric.py:176-180) */
#define MSS_ERR_INVALID_ARG (-1)
#define MSS_ERR_CUDA (-2)
#define MSS_ERR_NAN (-3)         /* sklearn: ValueError("Input contains NaN.") */
#define MSS_ERR_INF (-4)         /* sklearn: ValueError("Input contains infinity ...") */
#define MSS_ERR_WORKSPACE (-5)   /* workspace / capacity too small */
#define MSS_ERR_UNSUPPORTED (-6)

/* label element types accepted wherever a `labels` pointer appears */
#define MSS_LABEL_U8 1
#define MSS_LABEL_I32 4
#define MSS_LABEL_I64 8          /* what the reference testers pass (target.long(), test_deeplab.py:90) */

/* which score maps mss_deeplab_score computes (bit mask); higher = more anomalous for all four */
#define MSS_SCORE_ENERGY 1u      /* -(logsumexp_c x)            deepv3.py:251-253 */
#define MSS_SCORE_MAXLOGIT 2u    /* -max_c x                    north_star extra score */
#define MSS_SCORE_MSP 4u         /* 1 - max_c softmax(x)        north_star extra score */
#define MSS_SCORE_ENTROPY 8u     /* -sum_c p_c log p_c          north_star extra score */

MSS_API int mss_abi_version(void);
MSS_API const char *mss_last_error(void);
/* number of kernels this library has launched in the calling process (all threads); bench.py's gpu_launches */
MSS_API int64_t mss_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Streaming evaluator buffers (device side of the tester loop, test_deeplab.py:84-101 /
 * test_m2f.py:125-144: replaces per-batch .cpu().numpy() + np.concatenate with an on-device
 * append of order-preserving keys for the valid pixels only).
 * ------------------------------------------------------------------------------------------- */
#define MSS_EVAL_STATE_BYTES 64
/* Two key-only streams in ONE key buffer (ABI 2; ABI 1 stored (key, u8 label) pairs): the 0/1 label of a valid pixel
 * is the stream its key lives in, so no label byte is stored, sorted or exchanged between GPUs. */
typedef struct mss_eval_buffers {
    uint32_t *keys;     /* [capacity] ascending key order == descending float32 score; -0.0 == +0.0.
                         * in-distribution keys fill keys[0, n_neg) upwards,
                         * OOD keys fill keys[capacity - n_pos, capacity) downwards */
    void *state;        /* MSS_EVAL_STATE_BYTES device bytes: {u64 n_neg, u64 n_pos, u32 nan, u32 inf, u64 dropped, ...} */
    int64_t capacity;   /* in elements; n_neg + n_pos <= capacity */
} mss_eval_buffers;

/* zero the state (count = 0) */
MSS_API int mss_eval_reset(const mss_eval_buffers *ev, void *stream);
/* metric.py:171-172 selection (label == id_in / id_out, everything else ignored) + key build; appends
 * the valid pixels of (scores, labels)[0..n) to ev (order inside a stream is unspecified). */
MSS_API int mss_eval_append(const float *scores, const void *labels, int label_dtype, int64_t n,
                    int64_t id_in, int64_t id_out, const mss_eval_buffers *ev, void *stream);
/* read back {count = n_neg + n_pos, n_pos, nan_flag, inf_flag} (synchronises the stream);
 * MSS_ERR_WORKSPACE when the capacity was exceeded */
MSS_API int mss_eval_state_host(const mss_eval_buffers *ev, int64_t out_host[4], void *stream);

/* ---------------------------------------------------------------------------------------------
 * (a1, a2) DeepLabv3+ fused per-pixel scoring over NCHW fp32 logits [B, C, HW].
 * Replaces DeepWV3Plus.energy_func (lib/network/deepv3/deepv3.py:251-253) and adds the three
 * north_star scores.  Output pointers may be NULL for maps not selected in `which`.
 * If `ev` is non-NULL, `labels` ([B*HW], label_dtype) must be given: the score selected by
 * `key_which` (exactly one bit) is also appended to the evaluator for pixels with label
 * id_in / id_out -- the ignore-label masking of metric.py:171-172 fused into the scoring pass.
 * ------------------------------------------------------------------------------------------- */
MSS_API int mss_deeplab_score(const float *logits, int64_t B, int C, int64_t HW, unsigned which,
                      float *energy, float *maxlogit, float *msp, float *entropy,
                      const void *labels, int label_dtype, int64_t id_in, int64_t id_out,
                      unsigned key_which, const mss_eval_buffers *ev, void *stream);

/* (a3, a4) bilinear resize of [NC, h, w] -> [NC, H, W] fp32.  align_corners=1 replaces mynn.Upsample
 * (lib/network/deepv3/mynn.py:28-33); align_corners=0 replaces the F.interpolate calls at
 * lib/network/mask2former/maskformer_model.py:264-277. */
MSS_API int mss_upsample_bilinear(const float *in, int64_t NC, int h, int w, float *out, int H, int W,
                          int align_corners, void *stream);

/* deepv3.py:282-283 in one call: energy at head resolution [B,C,h,w] -> upsample (align_corners=True)
 * -> anomaly score [B,H,W].  `scratch` holds B*h*w floats. */
MSS_API int mss_deeplab_anomaly_score(const float *ood_logits, int64_t B, int C, int h, int w,
                              float *scratch, float *score, int H, int W, void *stream);

/* (DeepLab half) backward of the scoring ops, for the trainer's use of the same path with autograd
 * (train_deeplab.py:197-198: the loss of lib/loss.py:34-147 is a function of the anomaly-score map):
 *   mss_deeplab_energy_backward          grad_logits[b,c,p] = -softmax(logits[b,:,p])_c * grad_score[b,p]   (deepv3.py:251-253)
 *   mss_upsample_bilinear_backward       grad_in = A^T grad_out, A = the tap matrix of mss_upsample_bilinear (mynn.py:28-33)
 *   mss_deeplab_anomaly_score_backward   both fused (deepv3.py:283): grad_score [B,H,W] -> grad of ood_logits [B,C,h,w]
 * Gather form, no atomics: deterministic. */
MSS_API int mss_deeplab_energy_backward(const float *logits, const float *grad_score, int64_t B, int C, int64_t HW,
                                float *grad_logits, void *stream);
MSS_API int mss_upsample_bilinear_backward(const float *grad_out, int64_t NC, int h, int w, float *grad_in, int H, int W,
                                   int align_corners, void *stream);
MSS_API int mss_deeplab_anomaly_score_backward(const float *ood_logits, const float *grad_score, int64_t B, int C, int h,
                                       int w, int H, int W, float *grad_logits, void *stream);

/*  DeepLabv3+ head fusion, deepv3.py:279-283: the two bias-free 1x1 convolutions on the decoder
 * feature map and the energy score in one pass (tcgen05 3xTF32 GEMM, feature read once):
 *   feature [B, K, hw] NCHW fp32;  w_cls = final[-1].weight, w_ood = ood_head.weight, each [C, K] row-major
 *   dec1 [B, C, hw] = conv1x1(feature, w_cls)      (deepv3.py:279)   or NULL
 *   dec2 [B, C, hw] = conv1x1(feature, w_ood)      (deepv3.py:282)   or NULL
 *   energy [B, hw]  = -logsumexp_c dec2            (deepv3.py:251-253, before the Upsample of :283)   or NULL
 * Supported: C <= 24, K a multiple of 32 up to 256 (the model: C = 19, K = 256); otherwise
 * MSS_ERR_UNSUPPORTED.  workspace: mss_deeplab_head_workspace_bytes(K). */
MSS_API size_t mss_deeplab_head_workspace_bytes(int K);
MSS_API int mss_deeplab_head(const float *feature, int64_t B, int K, int64_t hw, const float *w_cls,
                     const float *w_ood, int C, float *dec1, float *dec2, float *energy, void *workspace,
                     size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (a4-a7) Mask2Former fused post-head inference.
 *   cls_logits  [B, Q, C+1]        pred_logits / pred_logits_ood
 *   mask_logits [B, Q, h, w]       pred_masks / pred_masks_ood at decoder resolution
 * computes, without ever materialising the [Q, Hp, Wp] tensor,
 *   P = softmax(cls)[..., :C]; S = sigmoid(bilinear(mask_logits -> Hp x Wp, align_corners=False));
 *   semseg[b, c, y, x] = sum_q P[b,q,c] S[b,q,y,x]                     (maskformer_model.py:343-345)
 *   extra[b][k]        = score_k * S[b, q_k]   for the kept queries     (maskformer_model.py:346-352)
 *   anomaly[b, y, x]   = 1 - max_c semseg[b, c, y, x]                   (train_m2f.py:402-407)
 * cropped to y < Hc, x < Wc (sem_seg_postprocess crop, maskformer_model.py:299-300; train_m2f.py:406).
 * Hp == h and Wp == w is the "already upsampled" case (drop-in semantic_inference / get_anomaly_score).
 * Any of semseg / anomaly may be NULL (not both).  Extra channels: keep_idx [B, Q] int32 lists, per
 * image, the kept query indices (first keep_count[b] entries valid; keep_count is a DEVICE int32 [B]) and
 * keep_score [B, Q] holds softmax(cls).max(-1) per query; plane k of image b goes to
 * extra + b*extra_batch_stride + k*Hc*Wc.  Pass NULL to skip.  semseg_batch_stride = elements between
 * consecutive images' semseg blocks (>= C*Hc*Wc), so the caller can place the extra planes right behind
 * the C planes of the same image (the torch.cat layout of maskformer_model.py:352).
 * workspace: mss_m2f_workspace_bytes(B, Q, C) device bytes.  flags: MSS_M2F_FORCE_GENERIC skips the
 * TMA-staged x4 kernel (testing).  Limits: Q <= 128, C <= 32; the fast path needs C == 19, Hp == 4h,
 * Wp == 4w, w % 4 == 0 (always true for the model: stride 4, size divisibility 32).
 * ------------------------------------------------------------------------------------------- */
#define MSS_M2F_FORCE_GENERIC 1u   /* any-resize kernel (no TMA) */
#define MSS_M2F_FORCE_FFMA 2u      /* TMA-staged x4 kernel with the FP32-FMA contraction instead of the 3xTF32 tensor-core one */
#define MSS_M2F_FORCE_MMASYNC 4u   /* 3xTF32 contraction through mma.sync (legacy tensor path) instead of tcgen05 + TMEM */
#define MSS_M2F_FORCE_TC5_PIXEL 8u /* tcgen05 kernel with one pixel per thread (first form) instead of the shared-tap 4x1 blocks */
MSS_API size_t mss_m2f_workspace_bytes(int64_t B, int Q, int C);
MSS_API int mss_m2f_semantic_inference(const float *cls_logits, const float *mask_logits,
                               int64_t B, int Q, int C, int h, int w, int Hp, int Wp, int Hc, int Wc,
                               float *semseg, int64_t semseg_batch_stride, float *anomaly,
                               const int32_t *keep_idx, const float *keep_score,
                               const int32_t *keep_count, float *extra, int64_t extra_batch_stride,
                               void *workspace, size_t workspace_bytes, unsigned flags, void *stream);

/* (Mask2Former half) backward of the fused anomaly score of mss_m2f_semantic_inference, for the
 * trainer's use of the same path with autograd (train_m2f.py:443 calls get_anomaly_score :387-407 inside the step; its
 * result feeds criterion.loss_ood, modeling/criterion.py:128-187, and lib/loss.py:119-147):
 *   grad_score [B, Hc, Wc] = dL/d(1 - max_c sum_q softmax(cls)[q,c] sigmoid(upsample(mask_logits))[q,px])
 *   grad_cls   [B, Q, C+1] = dL/dcls_logits     grad_masks [B, Q, h, w] = dL/dmask_logits (decoder resolution)
 * torch.max semantics (gradient to the first maximal class), exact adjoint of the align_corners=False upsample, no
 * atomics: deterministic.  The [Q, Hp, Wp] tensors are never materialised.  Q <= 128, C < 32.
 * workspace: mss_m2f_anomaly_backward_workspace_bytes(B, Q, C, Hc, Wc). */
MSS_API size_t mss_m2f_anomaly_backward_workspace_bytes(int64_t B, int Q, int C, int Hc, int Wc);
MSS_API int mss_m2f_anomaly_backward(const float *cls_logits, const float *mask_logits, const float *grad_score, int64_t B,
                             int Q, int C, int h, int w, int Hp, int Wp, int Hc, int Wc, float *grad_cls,
                             float *grad_masks, void *workspace, size_t workspace_bytes, void *stream);

/* (Mask2Former half) the mask-logit contraction in front of the scoring path,
 * lib/network/mask2former/modeling/transformer_decoder/mask2former_transformer_decoder.py:528-529 (pred_masks)
 * and :548-549 (pred_masks_ood):   outputs_mask = torch.einsum("bqc,bchw->bqhw", mask_embed, mask_features)
 *   mask_embed    [B, Q, K]   per-image query embeddings (output of the mask_embed MLP)
 *   mask_features [B, K, hw]  pixel-decoder features, NCHW fp32, read once
 *   mask_logits   [B, Q, hw]  = the decoder-resolution masks mss_m2f_semantic_inference takes
 * tcgen05 3xTF32 GEMM (fp32-level accuracy), one CTA per SM holding the image's pre-split embedding table.
 * Supported: Q <= 112, K a multiple of 32 up to 256, hw <= 2^31 / 256 (the model: Q = 100, K = 256); otherwise
 * MSS_ERR_UNSUPPORTED.  workspace: mss_m2f_mask_logits_workspace_bytes(B, K). */
MSS_API size_t mss_m2f_mask_logits_workspace_bytes(int64_t B, int K);
MSS_API int mss_m2f_mask_logits(const float *mask_embed, const float *mask_features, int64_t B, int Q, int K,
                        int64_t hw, float *mask_logits, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (a9-a13) exact, tie-aware AUROC / AP / FPR@95TPR.
 * Replaces eval_ood_measure (lib/utils/metric.py:170-180) and everything below it
 * (get_measures :130-153, sklearn roc_auc_score / average_precision_score at :142/:146,
 * fpr_and_fdr_at_recall :87-127).  Results are bit-identical to that computation.
 * ------------------------------------------------------------------------------------------- */
/* bytes of device workspace for n candidate pixels (one-shot) or n stored keys (from evaluator) */
MSS_API size_t mss_ood_metrics_workspace_bytes(int64_t n);

/* one-shot: scores [n] fp32 + labels [n]  ->  out_host = {AUROC, AP, FPR95}, counts_host = {P, N, T, T_roc}.
 * Returns MSS_EMPTY_CLASS when there is no label==id_in or no label==id_out pixel, MSS_ERR_NAN /
 * MSS_ERR_INF when a valid score is not finite.  Synchronises `stream`. */
MSS_API int mss_ood_metrics(const float *scores, const void *labels, int label_dtype, int64_t n,
                    int64_t id_in, int64_t id_out, void *workspace, size_t workspace_bytes,
                    double out_host[3], int64_t counts_host[4], void *stream);

/* same, from the keys accumulated in an evaluator (both streams are sorted in place).
 * workspace >= mss_ood_metrics_from_eval_workspace_bytes(count), count = mss_eval_state_host()[0]. */
MSS_API size_t mss_ood_metrics_from_eval_workspace_bytes(int64_t count);
MSS_API int mss_ood_metrics_from_eval(const mss_eval_buffers *ev, void *workspace, size_t workspace_bytes,
                              double out_host[3], int64_t counts_host[4], void *stream);

/* The multi-GPU form of the same: one process or thread per GPU, each holding the keys of ITS images in an
 * evaluator; every rank calls this with the same NCCL communicator and gets the metrics of the WHOLE dataset,
 * bit-identical to the single-GPU result (key-range exchange: all-reduce of a sampled key histogram -> splitters, local
 * partition + ncclSend/ncclRecv, local sort + merge-path counts with integer prefixes, all-gather of the per-threshold
 * counts, identical float64 tail).  nccl_comm is an ncclComm_t; NCCL is resolved at run time from the caller's process
 * (the library is not linked against it): MSS_ERR_UNSUPPORTED if it is not loaded.  Temporaries are taken from and
 * returned to the stream-ordered allocator (cudaMallocAsync) inside the call.  The evaluator's keys are left untouched.
 * Returns like mss_ood_metrics (MSS_EMPTY_CLASS / MSS_ERR_NAN / MSS_ERR_INF decided on the GLOBAL dataset). */
MSS_API int mss_ood_metrics_dist(const mss_eval_buffers *ev, void *nccl_comm, int rank, int world, double out_host[3],
                         int64_t counts_host[4], void *stream);

/* ---- stage-level entry points (used by the multi-GPU evaluator, which interleaves them with
 * collectives issued through torch.distributed) ------------------------------------------------ */
/* stable LSD radix sort (onesweep, key-only) of two independent key arrays, ascending, in place, by one sequence
 * of launches (the second array may be NULL / empty).  workspace: mss_sort_keys_workspace_bytes(n_a + n_b). */
MSS_API size_t mss_sort_keys_workspace_bytes(int64_t n_total);
MSS_API int mss_sort_keys(uint32_t *keys_a, int64_t n_a, uint32_t *keys_b, int64_t n_b, void *workspace,
                  size_t workspace_bytes, void *stream);
/* the same for the two streams of an evaluator; the stream sizes are read from the DEVICE state (no host round
 * trip), n_upper >= n_neg + n_pos sizes the grids.  workspace: mss_sort_keys_workspace_bytes(n_upper). */
MSS_API int mss_eval_sort(const mss_eval_buffers *ev, int64_t n_upper, void *workspace, size_t workspace_bytes,
                  void *stream);
/* histogram of the top `bits` (<= 16) key bits: hist[1 << bits] int64, overwritten (splitter selection) */
MSS_API int mss_keys_histogram(const uint32_t *keys, int64_t n, int bits, int64_t *hist, void *stream);
/* same over a systematic sample: only every `every`-th group of 4 consecutive keys is counted (every >= 1).
 * Splitters only have to be IDENTICAL on all ranks and roughly balanced -- exactness never depends on them --
 * so the multi-GPU evaluator histograms ~2^24 keys per rank instead of all of them. */
MSS_API int mss_keys_histogram_sampled(const uint32_t *keys, int64_t n, int bits, int every, int64_t *hist,
                               void *stream);
/* second level of the same (same sampling): hist[65536] of the LOW 16 bits of the keys whose top 16 bits equal prefix16.
 * Lets a splitter fall INSIDE a heavy top-16-bit bin (saturated or narrow-range scores), so that one rank does not
 * end up owning the whole dataset. */
MSS_API int mss_keys_histogram_refine(const uint32_t *keys, int64_t n, unsigned prefix16, int every, int64_t *hist,
                              void *stream);
/* stable partition of keys into `parts` (<= 256) destination ranges:
 * dest(key) = #{ j : key >= splitters[j] }, splitters ascending device array [parts-1].
 * out_counts_host[parts] receives the bucket sizes (synchronises the stream). */
MSS_API size_t mss_partition_workspace_bytes(int64_t n, int parts);
MSS_API int mss_partition_keys(const uint32_t *keys, int64_t n, const uint32_t *splitters, int parts,
                       uint32_t *keys_out, int64_t *out_counts_host, void *workspace, size_t workspace_bytes,
                       void *stream);
/* bucket sizes only (same dest rule as mss_partition_keys); workspace >= 2048 bytes; synchronises the stream */
MSS_API int mss_partition_count(const uint32_t *keys, int64_t n, const uint32_t *splitters, int parts,
                        int64_t *out_counts_host, void *workspace, size_t workspace_bytes, void *stream);
/* Fused partition + exchange: the same stable partition, but bucket d is stored straight into ITS OWN buffer --
 * typically the receive buffer of GPU d, mapped into this process (peer / symmetric memory), so the stores travel
 * over NVLink and no NCCL all-to-all and no local staging copy is needed.  Only 4-byte key stores cross the link.
 *   dst_keys_host[d]     device address (as an integer) of bucket d's uint32 buffer
 *   dst_offsets_host[d]  element offset inside it where this rank's block starts
 * The caller sizes the blocks with mss_partition_count (+ an all-gather across ranks) and must order this call
 * against the peers' use of the buffers (barrier before and after).  workspace: mss_partition_workspace_bytes(n, parts).
 * Synchronises the stream. */
MSS_API int mss_partition_scatter_keys(const uint32_t *keys, int64_t n, const uint32_t *splitters, int parts,
                               const uint64_t *dst_keys_host, const int64_t *dst_offsets_host,
                               void *workspace, size_t workspace_bytes, void *stream);
/* The same two steps for BOTH streams of an evaluator in one call each (one host synchronisation per step; the splitters
 * are a HOST array here).  n_neg / n_pos as read with mss_eval_state_host.
 *   mss_eval_partition_count    out_counts_host[0 .. parts) = in-distribution keys per destination, [parts .. 2 parts) = OOD
 *   mss_eval_partition_scatter  destination d's buffer dst_keys_host[d] receives this rank's in-distribution bucket at element
 *                               offset dst_neg_offsets_host[d] and its OOD bucket at dst_pos_offsets_host[d]
 * workspace: mss_eval_partition_workspace_bytes(n_neg, n_pos, parts). */
MSS_API size_t mss_eval_partition_workspace_bytes(int64_t n_neg, int64_t n_pos, int parts);
MSS_API int mss_eval_partition_count(const mss_eval_buffers *ev, int64_t n_neg, int64_t n_pos, const uint32_t *splitters_host,
                             int parts, int64_t *out_counts_host, void *workspace, size_t workspace_bytes, void *stream);
MSS_API int mss_eval_partition_scatter(const mss_eval_buffers *ev, int64_t n_neg, int64_t n_pos, const uint32_t *splitters_host,
                               int parts, const uint64_t *dst_keys_host, const int64_t *dst_neg_offsets_host,
                               const int64_t *dst_pos_offsets_host, void *workspace, size_t workspace_bytes, void *stream);
/* Exchange by REMOTE APPEND (the default multi-GPU path): the receive side is itself an evaluator buffer in peer-mapped
 * memory.  Every key of this rank's evaluator is appended to the evaluator of the rank that owns its key range: a tile
 * orders its keys by destination in shared memory, reserves its run with one system-scope atomicAdd on the destination's
 * stream counter (over NVLink) and stores it with one bulk copy.  No counting pass, no all-gather of bucket sizes, no
 * look-back, both streams in one launch.
 *   dst_keys_host[d], dst_state_host[d]  device addresses of rank d's key buffer / MSS_EVAL_STATE_BYTES state
 *   dst_capacity                         keys per destination buffer
 * The caller zeroes the destination states, orders the call between two barriers, and every rank then reads its own state
 * with mss_eval_state_host (which reports an overflow).  parts <= 32; workspace >= 4096 bytes.  Synchronises the stream. */
MSS_API int mss_eval_exchange_append(const mss_eval_buffers *ev, int64_t n_neg, int64_t n_pos, const uint32_t *splitters_host,
                             int parts, const uint64_t *dst_keys_host, const uint64_t *dst_state_host,
                             int64_t dst_capacity, void *workspace, size_t workspace_bytes, void *stream);
/* Streaming form of the same exchange: ENQUEUES (no host synchronisation) the exchange of a STAGING evaluator -- its
 * stream sizes are read from its device state, so a batch can be exchanged right behind the kernel that scored it while
 * the next batch is scored on another stream -- followed by `accum_state += staging state; staging state = 0`.
 * splitters_dev is a DEVICE array [parts - 1]; accum_state MSS_EVAL_STATE_BYTES device bytes.
 * workspace == NULL: every tile reserves its runs itself (one remote atomic per tile and destination);
 * workspace of mss_eval_exchange_stream_workspace_bytes(staging capacity, parts) bytes and parts <= 16: the batch is
 * counted per destination, ONE run per destination and stream is reserved remotely, and the keys are scattered with a
 * look-back between tiles (2 x parts remote atomics per batch). */
MSS_API size_t mss_eval_exchange_stream_workspace_bytes(int64_t staging_capacity, int parts);
MSS_API int mss_eval_exchange_stream(const mss_eval_buffers *staging, const uint32_t *splitters_dev, int parts,
                             const uint64_t *dst_keys_host, const uint64_t *dst_state_host, int64_t dst_capacity,
                             void *accum_state, void *workspace, size_t workspace_bytes, void *stream);
/* Staged form (the copy engines move the keys): ENQUEUES, on `stream`,
 *   1. the partition of a STAGING evaluator (sizes read from its device state) into `parts` LOCAL outbox buffers -- each an
 *      evaluator layout of outbox_capacity keys (negatives up from key 0, positives down from the end) with its own
 *      MSS_EVAL_STATE_BYTES state, zeroed here first;
 *   2. ONE reservation per destination and stream in the owners' receive states (recv_state_host[d]: device address of
 *      rank d's state, peer-mapped; system-scope atomicAdd) and the copy plan, written to `plan` (uint64[4 * parts], pinned
 *      host memory or device memory):  plan[4d + 0], [4d + 1] = number of in-distribution keys for rank d and the key index
 *      in rank d's receive buffer where that run starts;  [4d + 2], [4d + 3] the same for the OOD keys (the run occupies
 *      the LAST plan[4d + 2] keys of outbox d).  A run that does not fit is dropped and the owner's overflow is raised;
 *   3. accum_state += staging state; staging state = 0.
 * After the work enqueued here has completed (event), the caller copies run d from outbox d to the owner's key buffer at
 * the planned index -- e.g. mss_memcpy_async on another stream -- and may reuse the outboxes once those copies are done.
 * This is what replaces test_deeplab.py:94-101's per-batch .cpu().numpy() + the final np.concatenate across GPUs. */
MSS_API int mss_eval_exchange_stage(const mss_eval_buffers *staging, const uint32_t *splitters_dev, int parts,
                            const uint64_t *outbox_keys_host, const uint64_t *outbox_state_host, int64_t outbox_capacity,
                            const uint64_t *recv_state_host, int64_t recv_capacity, void *accum_state, uint64_t *plan,
                            void *stream);
/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream): device-to-device (also peer-mapped) block copy on the copy
 * engines, for hosts that have no CUDA runtime binding of their own */
MSS_API int mss_memcpy_async(void *dst, const void *src, size_t bytes, void *stream);
/* two sorted key arrays (negatives = in-distribution, positives = OOD) -> per distinct key of their union the
 * cumulative counts  tps[k] = pos_before + #{positives with key <= key_k},  fps[k] = neg_before + #{negatives with
 * key <= key_k}  (int64; one merge-path pass).  tps/fps need room for n_neg + n_pos entries.  *T_host = number of
 * distinct keys (synchronises the stream). */
MSS_API size_t mss_counts_workspace_bytes(int64_t n);
MSS_API int mss_counts_from_sorted(const uint32_t *neg_keys, int64_t n_neg, const uint32_t *pos_keys, int64_t n_pos,
                           int64_t pos_before, int64_t neg_before, int64_t *tps, int64_t *fps, int64_t *T_host,
                           void *workspace, size_t workspace_bytes, void *stream);
/* float64 tail over device int64 (tps, fps)[T]: out_host = {AUROC, AP, FPR95}; *T_roc_host = points kept
 * by roc_curve(drop_intermediate=True).  recall_level is 0.95 everywhere in the reference
 * (metric.py:130).  Replays numpy's pairwise summation tree exactly. */
MSS_API size_t mss_tail_workspace_bytes(int64_t T);
MSS_API int mss_metrics_tail(const int64_t *tps, const int64_t *fps, int64_t T, double recall_level,
                     void *workspace, size_t workspace_bytes, double out_host[3], int64_t *T_roc_host,
                     void *stream);

/* host-only helper (no device work): numpy's pairwise-summation tree (np.add.reduce over float64, leaves of
 * <= 128 terms, split at n/2 rounded down to a multiple of 8) that mss_metrics_tail replays for the AUROC /
 * AP sums of sklearn (_ranking.py auc / average_precision_score).  Returns leaf `leaf`'s [start, start+len)
 * and the number of leaves, computed by the same table descent the device kernels use. */
MSS_API int mss_pairwise_leaf_bounds(int64_t n, int64_t leaf, int64_t *start_host, int64_t *len_host,
                             int64_t *n_leaves_host);
/* host-only: np.sum(terms) over float64 through the same plan / descent / combine code (test hook) */
MSS_API int mss_pairwise_sum_host(const double *terms_host, int64_t n, double *out_host);

/* ---------------------------------------------------------------------------------------------
 *  mIoU half of lib/utils/metric.py: the confusion histogram of hist_info (:10-18),
 *   k = (gt >= 0) & (gt < n_cl); labeled = sum(k); correct = sum(pred[k] == gt[k]);
 *   hist = bincount(n_cl * gt[k] + pred[k], minlength = n_cl^2).reshape(n_cl, n_cl)
 * ACCUMULATED (+=) into caller-owned, caller-zeroed device int64 buffers, which is compute_metric's
 * `hist += d['hist']; correct += ...; labeled += ...` loop (:21-33) kept on the device:
 *   hist [n_cl * n_cl] int64 (row = gt, column = pred), labeled_correct [3] int64 = {labeled, correct,
 *   out-of-range count}.  n_cl <= 32.  A labeled pixel whose n_cl*gt + pred falls outside [0, n_cl^2)
 * makes numpy raise; here it is counted in labeled_correct[2] and reported (MSS_ERR_INVALID_ARG) by
 * mss_confusion_result.  The two update calls only enqueue work: a streaming evaluation synchronises once, at the end.
 * ------------------------------------------------------------------------------------------- */
/* pred, gt: class-index maps [n] of MSS_LABEL_* element types */
MSS_API int mss_confusion_hist(const void *pred, int pred_dtype, const void *gt, int gt_dtype, int64_t n,
                       int n_cl, int64_t *hist, int64_t *labeled_correct, void *stream);
/* fused: pred = argmax_c logits[b, c, p] (torch.argmax: first maximal index) over NCHW fp32 logits
 * [B, C, HW], never written to memory */
MSS_API int mss_confusion_from_logits(const float *logits, int64_t B, int C, int64_t HW, const void *gt,
                              int gt_dtype, int n_cl, int64_t *hist, int64_t *labeled_correct,
                              void *stream);

/* accumulators -> host (one D2H copy, synchronises): hist_host [n_cl * n_cl], labeled_correct_host = {labeled, correct} */
MSS_API int mss_confusion_result(const int64_t *hist, const int64_t *labeled_correct, int n_cl, int64_t *hist_host,
                         int64_t labeled_correct_host[2], void *stream);
/* host-only (no device work): compute_score (metric.py:42-49; per_class = 0) / compute_score_per_class (:51-64;
 * per_class = 1) on the float64 confusion accumulator of compute_metric (:22), operation by operation as numpy does it:
 *   iu_host [n_cl], class_acc_host [n_cl] (per_class only), out_host = {mean_IU, mean_IU_no_back | nan, mean_pixel_acc} */
MSS_API int mss_confusion_scores(const double *hist_host, int n_cl, double correct, double labeled, int per_class,
                         double *iu_host, double *class_acc_host, double out_host[3]);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer entry points: what a reference-side caller holding numpy / CPU tensors would call.
 * Inputs and outputs are HOST pointers (pinned memory recommended); the library stages them
 * through caller-provided device scratch in chunks, overlapping H2D, kernel and D2H on internal
 * streams forked from `stream`.  These are the `e2e` legs of bench.py.
 * ------------------------------------------------------------------------------------------- */
MSS_API size_t mss_deeplab_score_host_scratch_bytes(int64_t B, int C, int64_t HW, unsigned which);
MSS_API int mss_deeplab_score_host(const float *logits_host, int64_t B, int C, int64_t HW, unsigned which,
                           float *energy_host, float *maxlogit_host, float *msp_host, float *entropy_host,
                           void *device_scratch, size_t scratch_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MSS_B200_H */
