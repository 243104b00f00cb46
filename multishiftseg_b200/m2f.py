"""Mask2Former post-head inference: drop-ins for ``MaskFormer.semantic_inference``
(lib/network/mask2former/maskformer_model.py:341-354) and ``TrainM2FOOD.get_anomaly_score``
(train_m2f.py:387-407), plus the fused entry points that also absorb the ``F.interpolate`` calls of
``MaskFormer.forward`` (maskformer_model.py:264-277) and the ``sem_seg_postprocess`` crop (:299-300), so
that the [Q, H, W] mask tensor is never materialised.

All functions take and return CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import _lib as L

__all__ = ["semantic_inference", "get_anomaly_score", "post_head_inference", "anomaly_score_from_lowres", "mask_logits",
           "anomaly_score_from_features"]


def _keep_queries(mask_cls: torch.Tensor, num_classes: int):
    """maskformer_model.py:346-349, evaluated with the same torch ops on the tiny [B, Q, C+1] tensor so the
    boolean ``keep`` mask is bit-identical: labels != C, score > 0.95, 1 < label < 11."""
    scores, labels = F.softmax(mask_cls, dim=-1).max(-1)
    keep = labels.ne(num_classes) & (scores > 0.95) & (labels < 11) & (labels > 1)
    return scores, keep


def _run(cls_logits, mask_logits, padded_size, crop_size, want_semseg, want_anomaly, extra_channels, flags=0):
    L.require_cuda(cls_logits, "class logits")
    L.require_cuda(mask_logits, "mask logits")
    L.forbid_grad("the fused Mask2Former post-head kernel (semantic_inference / get_anomaly_score / post_head_inference)",
                  cls_logits, mask_logits)
    cls_logits = cls_logits.float().contiguous()
    mask_logits = mask_logits.float().contiguous()
    B, Q, C1 = cls_logits.shape
    Cn = C1 - 1
    Bm, Qm, h, w = mask_logits.shape
    if (Bm, Qm) != (B, Q):
        raise ValueError(f"class logits {tuple(cls_logits.shape)} and mask logits {tuple(mask_logits.shape)} disagree")
    Hp, Wp = int(padded_size[0]), int(padded_size[1])
    Hc, Wc = min(int(crop_size[0]), Hp), min(int(crop_size[1]), Wp)
    dev = cls_logits.device
    lib = L.load()
    semseg = anomaly = None
    keep_idx = keep_score = keep_count = None
    counts: List[int] = [0] * B
    if want_semseg:
        if extra_channels:
            scores, keep = _keep_queries(cls_logits, Cn)
            counts = keep.sum(-1).tolist()                       # K is data dependent (host sync, as in torch's boolean indexing)
            # stable order of kept queries: ascending q, exactly scores[keep] / mask_pred[keep]
            order = torch.argsort((~keep).to(torch.int8), dim=-1, stable=True).to(torch.int32).contiguous()
            keep_idx, keep_score = order, scores.contiguous()
            keep_count = keep.sum(-1).to(torch.int32).contiguous()
        kmax = max(counts) if counts else 0
        semseg = torch.empty((B, Cn + kmax, Hc, Wc), dtype=torch.float32, device=dev)
    if want_anomaly:
        anomaly = torch.empty((B, Hc, Wc), dtype=torch.float32, device=dev)
    nbytes = lib.mss_m2f_workspace_bytes(B, Q, Cn)
    ws = L.workspace(nbytes, dev)
    bstride = semseg.stride(0) if semseg is not None else 0
    extra_ptr = (semseg.data_ptr() + Cn * Hc * Wc * 4) if (semseg is not None and max(counts) > 0) else 0
    with torch.cuda.device(dev):
        rc = lib.mss_m2f_semantic_inference(
            cls_logits.data_ptr(), mask_logits.data_ptr(), B, Q, Cn, h, w, Hp, Wp, Hc, Wc,
            L.ptr(semseg), bstride, L.ptr(anomaly), L.ptr(keep_idx) if extra_ptr else 0,
            L.ptr(keep_score) if extra_ptr else 0, L.ptr(keep_count) if extra_ptr else 0, extra_ptr, bstride,
            ws.data_ptr(), nbytes, flags, L.stream_ptr(dev))
    L.check(rc, "mss_m2f_semantic_inference")
    return semseg, anomaly, counts


class _AnomalyFn(torch.autograd.Function):
    """1 - max_c sum_q softmax(cls)[q, c] sigmoid(upsample(masks))[q] with its backward (Mask2Former
    half): train_m2f.py:443 differentiates through get_anomaly_score (:387-407)."""

    @staticmethod
    def forward(ctx, cls_logits, mask_logits, Hp, Wp, Hc, Wc, flags):
        _, anomaly, _ = _run(cls_logits, mask_logits, (Hp, Wp), (Hc, Wc), False, True, False, flags)
        ctx.save_for_backward(cls_logits, mask_logits)
        ctx.sizes = (int(Hp), int(Wp), anomaly.shape[-2], anomaly.shape[-1])
        return anomaly

    @staticmethod
    def backward(ctx, grad):
        cls_logits, mask_logits = ctx.saved_tensors
        Hp, Wp, Hc, Wc = ctx.sizes
        cls = cls_logits.float().contiguous()
        masks = mask_logits.float().contiguous()
        g = grad.float().contiguous()
        B, Q, C1 = cls.shape
        h, w = masks.shape[-2:]
        lib = L.load()
        gc = torch.empty_like(cls)
        gm = torch.empty_like(masks)
        nbytes = lib.mss_m2f_anomaly_backward_workspace_bytes(B, Q, C1 - 1, Hc, Wc)
        ws = L.workspace(nbytes, cls.device)
        with torch.cuda.device(cls.device):
            rc = lib.mss_m2f_anomaly_backward(cls.data_ptr(), masks.data_ptr(), g.data_ptr(), B, Q, C1 - 1, h, w, Hp, Wp, Hc, Wc,
                                              gc.data_ptr(), gm.data_ptr(), ws.data_ptr(), nbytes, L.stream_ptr(cls.device))
        L.check(rc, "mss_m2f_anomaly_backward")
        return gc.to(cls_logits.dtype), gm.to(mask_logits.dtype), None, None, None, None, None


def _anomaly(cls_logits, mask_logits, padded_size, size, flags=0):
    if L.wants_grad(cls_logits, mask_logits):
        L.require_cuda(cls_logits, "class logits")
        L.require_cuda(mask_logits, "mask logits")
        return _AnomalyFn.apply(cls_logits, mask_logits, int(padded_size[0]), int(padded_size[1]), int(size[0]), int(size[1]),
                                flags)
    return _run(cls_logits, mask_logits, padded_size, size, False, True, False, flags)[1]


def semantic_inference(mask_cls: torch.Tensor, mask_pred: torch.Tensor, num_classes: Optional[int] = None) -> torch.Tensor:
    """maskformer_model.py:341-354 for ONE image: ``mask_cls`` [Q, C+1], ``mask_pred`` [Q, H, W] (already
    upsampled, as the reference passes it) -> [C + K, H, W]."""
    if num_classes is not None and num_classes != mask_cls.shape[-1] - 1:
        raise ValueError("num_classes must equal mask_cls.shape[-1] - 1")
    H, W = mask_pred.shape[-2:]
    semseg, _, counts = _run(mask_cls[None], mask_pred[None], (H, W), (H, W), True, False, True)
    return semseg[0, : mask_cls.shape[-1] - 1 + counts[0]]


def get_anomaly_score(other_outputs: Dict[str, torch.Tensor], size: Tuple[int, int]) -> torch.Tensor:
    """train_m2f.py:387-407: ``other_outputs["pred_masks_ood"]`` is the already-upsampled [B, Q, Hp, Wp].

    Differentiable: the reference also calls this inside the training step with autograd on (train_m2f.py:443); with
    inputs that require grad the result carries a grad_fn backed by ``mss_m2f_anomaly_backward`` (gradients w.r.t.
    ``pred_logits_ood`` and ``pred_masks_ood``).  The semantic-segmentation outputs (``semantic_inference``,
    ``post_head_inference``) stay forward-only and raise ``MssError`` for grad-tracked inputs."""
    cls = other_outputs["pred_logits_ood"]
    masks = other_outputs["pred_masks_ood"]
    Hp, Wp = masks.shape[-2:]
    return _anomaly(cls, masks, (Hp, Wp), size)


def post_head_inference(pred_logits: torch.Tensor, pred_masks: torch.Tensor, padded_size: Sequence[int],
                        image_sizes: Optional[Sequence[Sequence[int]]] = None, extra_channels: bool = True,
                        flags: int = 0) -> List[torch.Tensor]:
    """maskformer_model.py:264-300 fused: decoder-resolution ``pred_masks`` [B, Q, h, w] -> per image
    ``sem_seg`` [C + K_b, H_b, W_b] (upsample to ``padded_size``, semantic_inference, crop to the image size)."""
    B = pred_logits.shape[0]
    Cn = pred_logits.shape[-1] - 1
    if image_sizes is None:
        image_sizes = [tuple(padded_size)] * B
    sizes = {tuple(int(v) for v in s) for s in image_sizes}
    if len(sizes) == 1:
        semseg, _, counts = _run(pred_logits, pred_masks, padded_size, next(iter(sizes)), True, False, extra_channels, flags)
        return [semseg[b, : Cn + counts[b]] for b in range(B)]
    out = []
    for b in range(B):                                            # ragged image sizes: one launch per image
        s, _, counts = _run(pred_logits[b:b + 1], pred_masks[b:b + 1], padded_size, image_sizes[b], True, False,
                            extra_channels, flags)
        out.append(s[0, : Cn + counts[0]])
    return out


def anomaly_score_from_lowres(pred_logits_ood: torch.Tensor, pred_masks_ood: torch.Tensor, padded_size: Sequence[int],
                              size: Sequence[int], flags: int = 0) -> torch.Tensor:
    """maskformer_model.py:271-277 + train_m2f.py:387-407 fused: [B, Q, h, w] decoder masks -> [B, H, W] score.
    Differentiable w.r.t. both inputs (gradient at decoder resolution: the upsample's adjoint is fused in)."""
    return _anomaly(pred_logits_ood, pred_masks_ood, padded_size, size, flags)


def mask_logits(mask_embed: torch.Tensor, mask_features: torch.Tensor) -> torch.Tensor:
    """mask2former_transformer_decoder.py:529 / :549:
    ``outputs_mask = torch.einsum("bqc,bchw->bqhw", mask_embed, mask_features)`` -- ``mask_embed`` [B, Q, K],
    ``mask_features`` [B, K, h, w] -> [B, Q, h, w], on tcgen05 with 3xTF32 (fp32-level accuracy)."""
    L.require_cuda(mask_embed, "mask_embed")
    L.require_cuda(mask_features, "mask_features")
    L.forbid_grad("m2f.mask_logits", mask_embed, mask_features)
    if mask_embed.dim() != 3 or mask_features.dim() != 4:
        raise ValueError("mask_embed must be [B, Q, K] and mask_features [B, K, h, w]")
    e = mask_embed.float().contiguous()
    f = mask_features.float().contiguous()
    B, Q, K = e.shape
    Bf, Kf, h, w = f.shape
    if (Bf, Kf) != (B, K):
        raise ValueError(f"mask_embed {tuple(e.shape)} and mask_features {tuple(f.shape)} disagree")
    lib = L.load()
    out = torch.empty((B, Q, h, w), dtype=torch.float32, device=f.device)
    nbytes = lib.mss_m2f_mask_logits_workspace_bytes(B, K)
    ws = L.workspace(nbytes, f.device)
    with torch.cuda.device(f.device):
        rc = lib.mss_m2f_mask_logits(e.data_ptr(), f.data_ptr(), B, Q, K, h * w, out.data_ptr(), ws.data_ptr(), nbytes,
                                     L.stream_ptr(f.device))
    if rc == L.MSS_ERR_UNSUPPORTED:
        raise L.MssError("m2f.mask_logits: " + L.last_error())
    L.check(rc, "mss_m2f_mask_logits")
    return out


def anomaly_score_from_features(pred_logits_ood: torch.Tensor, mask_embed: torch.Tensor, mask_features: torch.Tensor,
                                padded_size: Sequence[int], size: Sequence[int]) -> torch.Tensor:
    """mask2former_transformer_decoder.py:548-549 + maskformer_model.py:271-277 + train_m2f.py:387-407 in two
    launches: mask-logit GEMM (decoder-resolution masks written once, 52 MB per 1024 x 2048 image), then the fused
    upsample / sigmoid / contraction / 1 - max kernel.  The [B, Q, H, W] tensors never exist."""
    return anomaly_score_from_lowres(pred_logits_ood, mask_logits(mask_embed, mask_features), padded_size, size)
