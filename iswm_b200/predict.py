"""Inference step of the hot path (reference: predict.py:258-290 `predict_mask`, batched).

    logits = model(x); prob = softmax(logits, 1); pred = prob[:, 1] > threshold; confidence = uint8(prob[:, 1] * 255)

For the reference's two-class models the final bilinear upsample, softmax, threshold and confidence map run as ONE
pass over the LOW-resolution logits (`iswm_predict_epilogue`) writing uint8 maps only - the full-resolution fp32
logits (268 MB at 8 x 2 x 2048^2) are never written or read; other class counts take `iswm_argmax_confusion` on the
upsampled logits. With `labels` given the same pass also accumulates the confusion matrix the way
evaluate_quantization.py:265-270 feeds `StreamMetrics`.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops


@torch.no_grad()
def predict_mask(model, images: torch.Tensor, threshold: float = 0.5, labels: Optional[torch.Tensor] = None,
                 metrics=None, fused: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """images: float32 [B,3,H,W] on CUDA (already normalised, predict.py:259-260). Returns (pred uint8 [B,H,W] in
    {0,1}, confidence uint8 [B,H,W]); `metrics` (a StreamMetrics) receives the confusion counts when labels are given."""
    if not images.is_cuda:
        raise RuntimeError("iswm_b200.predict runs on CUDA only (no CPU fallback)")
    net = getattr(model, "module", model)
    cm = metrics._cm() if (metrics is not None and labels is not None) else None
    H, W = images.shape[-2:]
    n_classes = net.engine().cls.cout
    if n_classes < 2:
        raise ValueError("predict_mask needs at least 2 classes (foreground = class 1)")
    if fused and n_classes == 2 and W % 4 == 0:
        # the reference's binary case: upsample + softmax + threshold + uint8 maps (+ confusion counts) in ONE kernel
        # over the low-resolution logits; bit-identical to the unfused path below
        lo = net.forward_lowres(images)
        _, pred, conf = ops.predict_epilogue(lo, H, W, labels if cm is not None else None, mode=1, threshold=threshold, out=cm)
        return pred, conf
    was_training = model.training
    model.eval()
    try:
        logits = model(images)
    finally:
        if was_training:
            model.train()
    _, pred, conf = ops.argmax_confusion(logits, labels if cm is not None else None, mode=1, threshold=threshold,
                                         want_pred=True, want_conf=True, out=cm)
    return pred, conf
