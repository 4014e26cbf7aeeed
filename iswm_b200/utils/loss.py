"""Criterion of the hot path.

The reference trains with `nn.CrossEntropyLoss(weight=class_weights, ignore_index=255,
reduction='mean')` (train.py:454-459); `CrossEntropyLoss` below has the same call signature and
semantics but runs as ONE fused CUDA pass (forward + gradient) fed by an integer class histogram.
`FocalLoss` / `create_loss` mirror utils/loss.py:14-39 (exported by the reference but unused).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn as nn

from .. import _lib, ops


class _WeightedCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, weight, ignore_index, hist_hook, hist_out):
        n_classes = logits.shape[1]
        hist = ops.class_hist(labels, n_classes)
        if hist_hook is not None:
            hist_hook(hist)                    # data parallel: SUM over ranks -> global-batch denominator
        if hist_out is not None:
            hist_out.append(hist)
        loss, grad = ops.wce_fwd_bwd(logits, labels, weight, hist, ignore_index, 1.0, logits.requires_grad)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g):
        grad = ctx.grad
        if grad is None:
            return (None,) * 6
        ctx.grad = None
        g = g.to(device=grad.device, dtype=torch.float32).contiguous()
        _lib.check(_lib.lib().iswm_scale_by_device_scalar(grad.data_ptr(), ops._FLOAT_CODE[grad.dtype], grad.numel(),
                                                          g.data_ptr(), ops._stream()), "scale_by_device_scalar")
        return grad, None, None, None, None, None


class CrossEntropyLoss(nn.Module):
    """criterion(logits float32|bf16 [B,C,H,W], labels int64|uint8 [B,H,W]) -> 0-dim float32 loss.

    Same semantics as torch.nn.CrossEntropyLoss(weight, ignore_index=255, reduction='mean'):
    L = sum_i w[y_i] * nll_i / sum_i w[y_i] over y_i != ignore_index; all-ignored batch -> nan.
    `hist_hook`, when set (iswm_b200.parallel), all-reduces the per-class pixel counts so that the
    normaliser is the GLOBAL batch's, as under the reference's nn.DataParallel (train.py:970,1046).
    """

    def __init__(self, weight: Optional[torch.Tensor] = None, ignore_index: int = 255, reduction: str = "mean",
                 check_labels: bool = False):
        """`check_labels`: labels outside [0, C) other than `ignore_index` are treated as ignored by the kernels (torch raises a
        device-side assert for them); with check_labels=True every call verifies the labels first and raises ValueError - a
        debugging aid that costs a pass over the labels and a host synchronisation."""
        super().__init__()
        self.check_labels = check_labels
        if reduction != "mean":
            raise NotImplementedError("the reference only uses reduction='mean' (train.py:457-459)")
        self.register_buffer("weight", None if weight is None else weight.detach().float().clone())
        self.ignore_index = ignore_index
        self.reduction = reduction
        self.hist_hook: Optional[Callable[[torch.Tensor], None]] = None
        self.last_hist: Optional[torch.Tensor] = None

    def forward(self, logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        if not logits.is_cuda:
            raise RuntimeError("iswm_b200 CrossEntropyLoss runs on CUDA only (no CPU fallback)")
        w = self.weight
        if w is not None and w.device != logits.device:
            w = w.to(logits.device)
        if self.check_labels:
            bad = ((labels < 0) | (labels >= logits.shape[1])) & (labels != self.ignore_index)
            n_bad = int(bad.sum())
            if n_bad:
                raise ValueError(f"CrossEntropyLoss: {n_bad} label values outside [0, {logits.shape[1]}) that are not ignore_index={self.ignore_index} "
                                 f"(first: {int(labels[bad][0])})")
        return _WeightedCEFn.apply(logits, labels, w, self.ignore_index, self.hist_hook, None)


class _FocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, weight, alpha, gamma, size_average, ignore_index):
        loss, grad = ops.focal_fwd_bwd(logits, labels, weight, alpha, gamma, size_average, ignore_index, logits.requires_grad)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g):
        grad = ctx.grad
        if grad is None:
            return (None,) * 7
        ctx.grad = None
        g = g.to(device=grad.device, dtype=torch.float32).contiguous()
        _lib.check(_lib.lib().iswm_scale_by_device_scalar(grad.data_ptr(), ops._FLOAT_CODE[grad.dtype], grad.numel(),
                                                          g.data_ptr(), ops._stream()), "scale_by_device_scalar")
        return grad, None, None, None, None, None, None


class FocalLoss(nn.Module):
    """utils/loss.py:14-35, same constructor and call signature; forward and gradient run as ONE fused CUDA pass
    (`iswm_focal_fwd_bwd`): focal_i = alpha * (1 - exp(-ce_i))^gamma * ce_i with ce_i the per-pixel weighted CE
    (0 where ignored), `.mean()` over ALL pixels when size_average else `.sum()`."""

    def __init__(self, alpha=1, gamma=0, size_average=True, ignore_index=255, weight=None):
        super().__init__()
        self.alpha, self.gamma, self.size_average, self.ignore_index = alpha, gamma, size_average, ignore_index
        self.register_buffer("weight", None if weight is None else weight.detach().float().clone())

    def forward(self, inputs, targets):
        if not inputs.is_cuda:
            raise RuntimeError("iswm_b200 FocalLoss runs on CUDA only (no CPU fallback)")
        if self.gamma < 0:
            raise ValueError("FocalLoss: gamma must be >= 0")
        return _FocalFn.apply(inputs, targets, self.weight, float(self.alpha), float(self.gamma), bool(self.size_average),
                              int(self.ignore_index))


def create_loss(loss_type="focal", temporal_loss="none", temporal_weight=0.5, **kwargs):
    """utils/loss.py:37-39."""
    return FocalLoss(**kwargs)
