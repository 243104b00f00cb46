"""CPU oracle (torch fp32) for the scoring half of the hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Parity status: PINNED for a1, a3, a5, a7 -- ``tests/golden/make_golden.py``
executes the reference's own function bodies (extracted by AST from the files
under /root/reference, because the modules themselves import detectron2 /
easydict which are absent) and the fixtures in ``tests/golden/`` hold their
outputs.  a4 (the ``F.interpolate`` call inlined in ``MaskFormer.forward``) is a
one-line torch call restated verbatim.  a2 (max-logit / max-softmax / entropy)
has NO reference code: "parity unpinned" -- defined by the expressions below.

Every function cites the reference lines it follows (relative to /root/reference).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F


# ---- DeepLabv3+ ----------------------------------------------------------------
def energy_func(logit: torch.Tensor) -> torch.Tensor:
    """lib/network/deepv3/deepv3.py:251-253."""
    return -(1.0 * torch.logsumexp(logit, dim=1))


def Upsample(x: torch.Tensor, size) -> torch.Tensor:
    """lib/network/deepv3/mynn.py:28-33 (bilinear, align_corners=True)."""
    return F.interpolate(x, size=size, mode="bilinear", align_corners=True)


def deeplab_anomaly_score(ood_logit: torch.Tensor, size) -> torch.Tensor:
    """lib/network/deepv3/deepv3.py:283 -- energy at head resolution, then upsample."""
    return Upsample(energy_func(ood_logit).unsqueeze(1), size).squeeze(1)


def deeplab_head(feature: torch.Tensor, w_cls: torch.Tensor, w_ood: torch.Tensor, size=None):
    """lib/network/deepv3/deepv3.py:279-283: ``dec1 = self.final[-1](feature)``,
    ``dec2 = self.ood_head(feature)`` -- both ``nn.Conv2d(256, num_classes, kernel_size=1, bias=False)``
    (deepv3.py:245-247), i.e. ``F.conv2d`` with a [C, K, 1, 1] weight -- then the energy of dec2 and its
    align_corners=True upsample.  Returns (dec1, anomaly_score, dec2).  Parity: the ops are torch's own; no
    reference-generated fixture (the lines sit inside ``forward``, which needs the whole network)."""
    dec1 = F.conv2d(feature, w_cls.reshape(w_cls.shape[0], -1, 1, 1))
    dec2 = F.conv2d(feature, w_ood.reshape(w_ood.shape[0], -1, 1, 1))
    e = energy_func(dec2)
    score = e if size is None else Upsample(e.unsqueeze(1), size).squeeze(1)
    return dec1, score, dec2


def maxlogit_score(logit: torch.Tensor) -> torch.Tensor:
    """north_star extra score (no reference code): -max_c x."""
    return -logit.max(dim=1)[0]


def msp_score(logit: torch.Tensor) -> torch.Tensor:
    """north_star extra score: 1 - max_c softmax(x)."""
    return 1.0 - torch.softmax(logit, dim=1).max(dim=1)[0]


def entropy_score(logit: torch.Tensor) -> torch.Tensor:
    """north_star extra score: -sum_c p_c log p_c."""
    logp = torch.log_softmax(logit, dim=1)
    return -(logp.exp() * logp).sum(dim=1)


# ---- Mask2Former ---------------------------------------------------------------
def upsample_masks(mask_pred: torch.Tensor, size) -> torch.Tensor:
    """lib/network/mask2former/maskformer_model.py:264-269 and :271-277."""
    return F.interpolate(mask_pred, size=size, mode="bilinear", align_corners=False)


def semantic_inference(mask_cls: torch.Tensor, mask_pred: torch.Tensor, num_classes: int = 19):
    """lib/network/mask2former/maskformer_model.py:341-354 (19 + K channels)."""
    mask_cls_f = F.softmax(mask_cls, dim=-1)[..., :-1]
    mask_pred_f = mask_pred.sigmoid()
    semseg = torch.einsum("qc,qhw->chw", mask_cls_f, mask_pred_f)
    scores, labels = F.softmax(mask_cls, dim=-1).max(-1)
    keep = labels.ne(num_classes) & (scores > 0.95) & (labels < 11) & (labels > 1)
    cur_prob_masks = scores[keep].view(-1, 1, 1) * mask_pred_f[keep]
    return torch.cat((semseg, cur_prob_masks), 0)


def get_anomaly_score(other_outputs: Dict[str, torch.Tensor], size: Tuple[int, int]) -> torch.Tensor:
    """train_m2f.py:387-407 (masks already upsampled by maskformer_model.py:271-277)."""
    class_probs = F.softmax(other_outputs["pred_logits_ood"], dim=-1)[..., :-1]
    mask_probs = other_outputs["pred_masks_ood"].sigmoid()
    u = torch.einsum("bqc,bqhw->bchw", class_probs, mask_probs)
    u = u[:, :, :size[0], :size[1]]
    return 1 - torch.max(u, dim=1)[0]


def m2f_post_head(pred_logits, pred_masks_lowres, padded_size, image_size, num_classes: int = 19):
    """maskformer_model.py:264-300 for one batch: upsample -> semantic_inference ->
    sem_seg_postprocess crop (identity resize).  Returns list of [19+K,H,W]."""
    up = upsample_masks(pred_masks_lowres, padded_size)
    out = []
    for cls, m in zip(pred_logits, up):
        r = semantic_inference(cls, m, num_classes)
        out.append(r[:, :image_size[0], :image_size[1]])
    return out


def m2f_anomaly_from_lowres(pred_logits_ood, pred_masks_ood_lowres, padded_size, image_size):
    """maskformer_model.py:271-277 + train_m2f.py:387-407 in sequence."""
    up = upsample_masks(pred_masks_ood_lowres, padded_size)
    return get_anomaly_score({"pred_logits_ood": pred_logits_ood, "pred_masks_ood": up}, image_size)


def m2f_mask_logits(mask_embed: torch.Tensor, mask_features: torch.Tensor) -> torch.Tensor:
    """lib/network/mask2former/modeling/transformer_decoder/mask2former_transformer_decoder.py:529 and :549
    (SURVEY 8f-1): ``outputs_mask = torch.einsum("bqc,bchw->bqhw", mask_embed, mask_features)``.  The op is
    torch's own; no reference-generated fixture (the line sits inside the decoder ``forward``)."""
    return torch.einsum("bqc,bchw->bqhw", mask_embed, mask_features)
