"""DeepLabv3+ scoring: drop-ins for ``DeepWV3Plus.energy_func`` (lib/network/deepv3/deepv3.py:251-253),
``mynn.Upsample`` (lib/network/deepv3/mynn.py:28-33) and the OOD-head tail of the forward
(deepv3.py:282-283), plus the fused multi-score entry point named by the north star
(energy / max-logit / max-softmax / entropy in one pass over the logits).

All functions take and return CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional, Sequence, Tuple

import torch

from . import _lib as L

__all__ = ["energy_func", "Upsample", "anomaly_score", "score_maps", "head_scores", "SCORES"]

SCORES = ("energy", "maxlogit", "msp", "entropy")


def _prep_logits(logit: torch.Tensor) -> torch.Tensor:
    L.require_cuda(logit, "logit")
    if logit.dim() < 2:
        raise ValueError("logit must be [B, C, ...]")
    if logit.dtype != torch.float32:
        logit = logit.float()
    return logit.contiguous()


def score_maps(logit: torch.Tensor, which: Iterable[str] = ("energy",), *, labels: Optional[torch.Tensor] = None,
               evaluator=None, key: Optional[str] = None, id_in: int = 0, id_out: int = 1) -> Dict[str, torch.Tensor]:
    """One pass over NCHW logits -> the requested score maps ``{name: [B, *spatial]}``.

    With ``evaluator`` (a ``metric.PairBuffer``) and ``labels`` ([B, *spatial], uint8/int32/int64) the map
    named ``key`` is also appended -- for pixels labelled ``id_in`` / ``id_out`` only -- to the on-device
    evaluator inside the same kernel (ignore-label masking fused into scoring).
    """
    L.forbid_grad("deeplab.score_maps", logit)     # (the autograd Functions below call it with grad mode off)
    logit = _prep_logits(logit)
    which = tuple(which)
    mask = 0
    for w in which:
        if w not in L.SCORE_BITS:
            raise ValueError(f"unknown score {w!r}; choose from {SCORES}")
        mask |= L.SCORE_BITS[w]
    if mask == 0:
        raise ValueError("no score selected")
    B, Cn = logit.shape[0], logit.shape[1]
    spatial = tuple(logit.shape[2:])
    HW = 1
    for s in spatial:
        HW *= s
    out = {w: torch.empty((B,) + spatial, dtype=torch.float32, device=logit.device) for w in which}
    ev_ref, lab_ptr, lab_code, key_bit = None, 0, 0, 0
    if evaluator is not None:
        if labels is None or key is None:
            raise ValueError("evaluator needs labels and key")
        L.require_cuda(labels, "labels")
        labels = labels.contiguous()
        if labels.numel() != B * HW:
            raise ValueError("labels must have one entry per pixel")
        if key not in which:
            raise ValueError("key must be one of the computed scores")
        ev_ref, lab_ptr, lab_code, key_bit = C.byref(evaluator.c), labels.data_ptr(), L.label_code(labels), L.SCORE_BITS[key]
    with torch.cuda.device(logit.device):
        rc = L.load().mss_deeplab_score(
            logit.data_ptr(), B, Cn, HW, mask, L.ptr(out.get("energy")), L.ptr(out.get("maxlogit")),
            L.ptr(out.get("msp")), L.ptr(out.get("entropy")), lab_ptr, lab_code, id_in, id_out, key_bit, ev_ref,
            L.stream_ptr(logit.device))
    L.check(rc, "mss_deeplab_score")
    return out


def _wants_grad(*ts: torch.Tensor) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)


class _EnergyFn(torch.autograd.Function):
    """energy_func with its backward (train_deeplab.py:197-198 differentiates through it)."""

    @staticmethod
    def forward(ctx, logit):
        x = _prep_logits(logit)
        ctx.save_for_backward(x)
        ctx.in_dtype = logit.dtype
        return score_maps(x, ("energy",))["energy"]

    @staticmethod
    def backward(ctx, grad):
        (x,) = ctx.saved_tensors
        g = grad.float().contiguous()
        B, Cn = x.shape[0], x.shape[1]
        HW = x.numel() // max(B * Cn, 1)
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            rc = L.load().mss_deeplab_energy_backward(x.data_ptr(), g.data_ptr(), B, Cn, HW, gx.data_ptr(),
                                                      L.stream_ptr(x.device))
        L.check(rc, "mss_deeplab_energy_backward")
        return gx.to(ctx.in_dtype)


class _UpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H, W, align_corners):
        ctx.shape, ctx.align, ctx.in_dtype = tuple(x.shape), bool(align_corners), x.dtype
        return _upsample_forward(x, H, W, align_corners)

    @staticmethod
    def backward(ctx, grad):
        N, Cn, h, w = ctx.shape
        g = grad.float().contiguous()
        H, W = g.shape[-2:]
        gx = torch.empty(ctx.shape, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            rc = L.load().mss_upsample_bilinear_backward(g.data_ptr(), N * Cn, h, w, gx.data_ptr(), H, W,
                                                         1 if ctx.align else 0, L.stream_ptr(g.device))
        L.check(rc, "mss_upsample_bilinear_backward")
        return gx.to(ctx.in_dtype), None, None, None


class _AnomalyScoreFn(torch.autograd.Function):
    """deepv3.py:283 with a fused backward: the upsampled gradient is gathered per head-resolution pixel and
    turned into -softmax(dec2) * g in one kernel."""

    @staticmethod
    def forward(ctx, ood_logit, H, W):
        x = _prep_logits(ood_logit)
        ctx.save_for_backward(x)
        ctx.size, ctx.in_dtype = (H, W), ood_logit.dtype
        return _anomaly_score_forward(x, H, W)

    @staticmethod
    def backward(ctx, grad):
        (x,) = ctx.saved_tensors
        B, Cn, h, w = x.shape
        H, W = ctx.size
        g = grad.float().contiguous()
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            rc = L.load().mss_deeplab_anomaly_score_backward(x.data_ptr(), g.data_ptr(), B, Cn, h, w, H, W, gx.data_ptr(),
                                                             L.stream_ptr(x.device))
        L.check(rc, "mss_deeplab_anomaly_score_backward")
        return gx.to(ctx.in_dtype), None, None


def energy_func(logit: torch.Tensor) -> torch.Tensor:
    """``-(1. * torch.logsumexp(logit, dim=1))`` -- deepv3.py:251-253 (call as a function, or bind it
    as the model's method: ``DeepWV3Plus.energy_func = lambda self, x: energy_func(x)``).  Differentiable: with
    autograd enabled and ``logit.requires_grad`` the backward (``-softmax(logit) * grad``) runs as a CUDA kernel."""
    if _wants_grad(logit):
        return _EnergyFn.apply(logit)
    return score_maps(logit, ("energy",))["energy"]


def Upsample(x: torch.Tensor, size: Sequence[int], align_corners: bool = True) -> torch.Tensor:
    """mynn.py:28-33: ``F.interpolate(x, size=size, mode='bilinear', align_corners=True)`` for [N,C,h,w]
    (differentiable: the backward is the exact adjoint of the forward's taps, gather form, deterministic)."""
    L.require_cuda(x, "x")
    if x.dim() != 4:
        raise ValueError("Upsample expects [N, C, h, w]")
    if _wants_grad(x):
        return _UpsampleFn.apply(x, int(size[0]), int(size[1]), align_corners)
    return _upsample_forward(x, int(size[0]), int(size[1]), align_corners)


def _upsample_forward(x: torch.Tensor, H: int, W: int, align_corners: bool) -> torch.Tensor:
    x = x.float().contiguous() if x.dtype != torch.float32 else x.contiguous()
    N, Cn, h, w = x.shape
    out = torch.empty((N, Cn, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.load().mss_upsample_bilinear(x.data_ptr(), N * Cn, h, w, out.data_ptr(), H, W,
                                            1 if align_corners else 0, L.stream_ptr(x.device))
    L.check(rc, "mss_upsample_bilinear")
    return out


def anomaly_score(ood_logit: torch.Tensor, size: Sequence[int]) -> torch.Tensor:
    """deepv3.py:283: ``Upsample(self.energy_func(dec2).unsqueeze(1), x_size[2:]).squeeze(1)`` (differentiable)."""
    if ood_logit.dim() != 4:
        raise ValueError("ood_logit must be [B, C, h, w]")
    if _wants_grad(ood_logit):
        L.require_cuda(ood_logit, "logit")
        return _AnomalyScoreFn.apply(ood_logit, int(size[0]), int(size[1]))
    return _anomaly_score_forward(_prep_logits(ood_logit), int(size[0]), int(size[1]))


def _anomaly_score_forward(ood_logit: torch.Tensor, H: int, W: int) -> torch.Tensor:
    B, Cn, h, w = ood_logit.shape
    scratch = torch.empty((B, h, w), dtype=torch.float32, device=ood_logit.device)
    out = torch.empty((B, H, W), dtype=torch.float32, device=ood_logit.device)
    with torch.cuda.device(ood_logit.device):
        rc = L.load().mss_deeplab_anomaly_score(ood_logit.data_ptr(), B, Cn, h, w, scratch.data_ptr(), out.data_ptr(),
                                                H, W, L.stream_ptr(ood_logit.device))
    L.check(rc, "mss_deeplab_anomaly_score")
    return out


def head_scores(feature: torch.Tensor, w_cls: torch.Tensor, w_ood: torch.Tensor, size: Optional[Sequence[int]] = None,
                want_dec2: bool = False):
    """deepv3.py:279-283 fused: ``feature`` [B, K, h, w] is read ONCE; returns

        dec1 = final[-1](feature)                       [B, C, h, w]   (feed ``Upsample(dec1, x_size[2:])`` for ``logit``)
        anomaly_score = Upsample(energy_func(ood_head(feature)).unsqueeze(1), size).squeeze(1)   [B, H, W]
                        (or the head-resolution energy [B, h, w] when ``size`` is None)
        dec2 (only with ``want_dec2``)

    ``w_cls`` / ``w_ood`` are the weights of the two bias-free 1x1 convolutions ([C, K] or [C, K, 1, 1]).
    Forward only: with autograd enabled and inputs that require grad it raises ``MssError`` (no backward kernel
    for the head GEMM) instead of returning tensors without grad_fn -- use it under ``torch.no_grad()``."""
    L.require_cuda(feature, "feature")
    L.forbid_grad("deeplab.head_scores (fused head GEMM)", feature, w_cls, w_ood)
    if feature.dim() != 4:
        raise ValueError("feature must be [B, K, h, w]")
    x = feature.float().contiguous()
    B, K, h, w = x.shape
    wc = L.require_cuda(w_cls, "w_cls").float().reshape(w_cls.shape[0], -1).contiguous()
    wo = L.require_cuda(w_ood, "w_ood").float().reshape(w_ood.shape[0], -1).contiguous()
    Cn = wc.shape[0]
    if wc.shape != (Cn, K) or wo.shape != (Cn, K):
        raise ValueError("w_cls and w_ood must both be [C, K] (or [C, K, 1, 1]) with K = feature channels")
    lib = L.load()
    dec1 = torch.empty((B, Cn, h, w), dtype=torch.float32, device=x.device)
    dec2 = torch.empty((B, Cn, h, w), dtype=torch.float32, device=x.device) if want_dec2 else None
    energy = torch.empty((B, h, w), dtype=torch.float32, device=x.device)
    nbytes = lib.mss_deeplab_head_workspace_bytes(K)
    ws = L.workspace(nbytes, x.device)
    with torch.cuda.device(x.device):
        rc = lib.mss_deeplab_head(x.data_ptr(), B, K, h * w, wc.data_ptr(), wo.data_ptr(), Cn, dec1.data_ptr(), L.ptr(dec2),
                                  energy.data_ptr(), ws.data_ptr(), nbytes, L.stream_ptr(x.device))
    if rc == L.MSS_ERR_UNSUPPORTED:
        raise L.MssError("deeplab.head_scores: " + L.last_error())
    L.check(rc, "mss_deeplab_head")
    score = energy if size is None else Upsample(energy.unsqueeze(1), size).squeeze(1)
    return (dec1, score, dec2) if want_dec2 else (dec1, score)


def score_maps_host(logits_host: torch.Tensor, which: Iterable[str] = ("energy",), scratch: Optional[torch.Tensor] = None,
                    out: Optional[Dict[str, torch.Tensor]] = None, device=None) -> Dict[str, torch.Tensor]:
    """Host-buffer variant (what a caller holding CPU tensors uses; bench.py's e2e leg): logits and the
    returned maps live in (preferably pinned) host memory, the library pipelines H2D / kernel / D2H."""
    if logits_host.is_cuda:
        raise ValueError("score_maps_host takes a CPU tensor; use score_maps for CUDA tensors")
    logits_host = logits_host.float().contiguous() if logits_host.dtype != torch.float32 else logits_host.contiguous()
    which = tuple(which)
    mask = 0
    for w in which:
        mask |= L.SCORE_BITS[w]
    B, Cn = logits_host.shape[0], logits_host.shape[1]
    spatial = tuple(logits_host.shape[2:])
    HW = 1
    for s in spatial:
        HW *= s
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lib = L.load()
    nbytes = lib.mss_deeplab_score_host_scratch_bytes(B, Cn, HW, mask)
    if scratch is None or scratch.numel() < nbytes:
        scratch = L.workspace(nbytes, dev)
    if out is None:
        out = {w: torch.empty((B,) + spatial, dtype=torch.float32, pin_memory=True) for w in which}
    with torch.cuda.device(dev):
        rc = lib.mss_deeplab_score_host(logits_host.data_ptr(), B, Cn, HW, mask, L.ptr(out.get("energy")),
                                        L.ptr(out.get("maxlogit")), L.ptr(out.get("msp")), L.ptr(out.get("entropy")),
                                        scratch.data_ptr(), scratch.numel(), L.stream_ptr(dev))
    L.check(rc, "mss_deeplab_score_host")
    return out
