"""Inference step of the hot path (reference: predict.py:258-290 `predict_mask`, batched).

    logits = model(x); prob = softmax(logits, 1); pred = prob[:, 1] > threshold; confidence = uint8(prob[:, 1] * 255)

The softmax / threshold / confidence map run as ONE pass over the logits (`iswm_argmax_confusion`, mode 1) writing
uint8 maps only; with `labels` given the same pass also accumulates the confusion matrix the way
evaluate_quantization.py:265-270 feeds `StreamMetrics`.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops


@torch.no_grad()
def predict_mask(model, images: torch.Tensor, threshold: float = 0.5, labels: Optional[torch.Tensor] = None,
                 metrics=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """images: float32 [B,3,H,W] on CUDA (already normalised, predict.py:259-260). Returns (pred uint8 [B,H,W] in
    {0,1}, confidence uint8 [B,H,W]); `metrics` (a StreamMetrics) receives the confusion counts when labels are given."""
    if not images.is_cuda:
        raise RuntimeError("iswm_b200.predict runs on CUDA only (no CPU fallback)")
    was_training = model.training
    model.eval()
    try:
        logits = model(images)
    finally:
        if was_training:
            model.train()
    if logits.shape[1] < 2:
        raise ValueError("predict_mask needs at least 2 classes (foreground = class 1)")
    cm = metrics._cm() if (metrics is not None and labels is not None) else None
    _, pred, conf = ops.argmax_confusion(logits, labels if cm is not None else None, mode=1, threshold=threshold,
                                         want_pred=True, want_conf=True, out=cm)
    return pred, conf
