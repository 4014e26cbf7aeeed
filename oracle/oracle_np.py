"""CPU oracle for the integer / loss part of the ISWM hot path — TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's algorithm, each function citing the reference
file:line it follows (paths relative to the reference tree). Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (iswm_b200/) never does.

Pinning: the reference ships no tests or golden vectors for this path,
so this oracle is pinned against outputs of the reference code itself, generated in the
build container by oracle/gen_golden.py (which imports /root/reference) and committed
under tests/golden/; tests/test_oracle_golden.py re-checks it on every run.
"""
from __future__ import annotations

import numpy as np


def class_pixel_counts(labels: np.ndarray):
    """train.py:401-402 — `(labels == 0).sum()`, `(labels == 1).sum()` as Python ints."""
    labels = np.asarray(labels)
    return int((labels == 0).sum()), int((labels == 1).sum())


def class_weights(black_pixels: int, white_pixels: int) -> np.ndarray:
    """train.py:404-410 — `[1.0, np.sqrt(black/white)]` stored as a FloatTensor (fp32)."""
    return np.array([1.0, np.sqrt(black_pixels / white_pixels)], dtype=np.float32)


def class_hist(labels: np.ndarray, n_classes: int) -> np.ndarray:
    """Per-class counts behind the weighted-mean denominator of nn.CrossEntropyLoss
    (train.py:457-459); generalises class_pixel_counts to n classes."""
    labels = np.asarray(labels).reshape(-1).astype(np.int64)
    ok = (labels >= 0) & (labels < n_classes)
    return np.bincount(labels[ok], minlength=n_classes).astype(np.int64)


def weighted_ce(logits: np.ndarray, labels: np.ndarray, weight=None, ignore_index: int = 255):
    """nn.CrossEntropyLoss(weight, ignore_index=255, reduction='mean') forward and its
    gradient w.r.t. logits (train.py:454-459 criterion; :1046 forward; :1048 backward).
    logits [B,C,...] ; labels [B,...]. Returns (loss float64, grad float64 like logits)."""
    x = np.asarray(logits, dtype=np.float64)
    y = np.asarray(labels).astype(np.int64)
    B, C = x.shape[0], x.shape[1]
    xm = np.moveaxis(x, 1, -1).reshape(-1, C)          # [N, C]
    yf = y.reshape(-1)
    w = np.ones(C, dtype=np.float64) if weight is None else np.asarray(weight, dtype=np.float64)
    valid = (yf != ignore_index) & (yf >= 0) & (yf < C)
    m = xm.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(xm - m).sum(axis=1))
    yc = np.where(valid, yf, 0)
    nll = lse - xm[np.arange(xm.shape[0]), yc]
    wy = np.where(valid, w[yc], 0.0)
    den = wy.sum()
    with np.errstate(invalid="ignore", divide="ignore"):
        loss = (wy * nll).sum() / den
        p = np.exp(xm - lse[:, None])
        onehot = np.zeros_like(p)
        onehot[np.arange(xm.shape[0]), yc] = 1.0
        g = wy[:, None] * (p - onehot) / den
    g[~valid] = 0.0
    g = np.moveaxis(g.reshape(x.shape[:1] + x.shape[2:] + (C,)), -1, 1)
    return float(loss), g


def fast_hist(label_true: np.ndarray, label_pred: np.ndarray, n_classes: int) -> np.ndarray:
    """metrics/stream_metrics.py:24-31 — `_fast_hist`: rows = true, cols = pred."""
    label_true = np.asarray(label_true).reshape(-1)
    label_pred = np.asarray(label_pred).reshape(-1)
    mask = (label_true >= 0) & (label_true < n_classes)
    return np.bincount(
        n_classes * label_true[mask].astype(int) + label_pred[mask].astype(int),
        minlength=n_classes ** 2,
    ).reshape(n_classes, n_classes)


def foreground_metrics(hist: np.ndarray, fg: int = 1):
    """metrics/stream_metrics.py:33-63 — (miou, fg_iou, precision, recall, f1), eps=1e-7."""
    hist = np.asarray(hist, dtype=np.float64)
    tp = hist[fg, fg]
    fp = hist[:, fg].sum() - tp
    fn = hist[fg, :].sum() - tp
    eps = 1e-7
    fg_iou = tp / (tp + fp + fn + eps)
    precision = tp / (tp + fp + eps)
    recall = tp / (tp + fn + eps)
    f1 = 2 * precision * recall / (precision + recall + eps)
    btp = hist[0, 0]
    bfp = hist[:, 0].sum() - btp
    bfn = hist[0, :].sum() - btp
    bg_iou = btp / (btp + bfp + bfn + eps)
    return (bg_iou + fg_iou) / 2.0, fg_iou, precision, recall, f1


def argmax_pred(logits: np.ndarray) -> np.ndarray:
    """train.py:644,659 — `logits.max(1)[1]` (first maximum wins)."""
    return np.argmax(np.asarray(logits), axis=1).astype(np.int64)


def threshold_pred(logits: np.ndarray, threshold: float = 0.5):
    """predict.py:264-278 — softmax over dim 1, foreground prob > threshold; also the
    `uint8(prob*255)` confidence map of predict.py:285-288. fp32 like the reference."""
    x = np.asarray(logits, dtype=np.float32)
    m = x.max(axis=1, keepdims=True)
    e = np.exp(x - m, dtype=np.float32)
    p1 = (e[:, 1] / e.sum(axis=1, dtype=np.float32)).astype(np.float32)
    return (p1 > np.float32(threshold)).astype(np.int64), (p1 * np.float32(255.0)).astype(np.uint8)
