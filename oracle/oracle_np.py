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


def focal_loss(logits: np.ndarray, labels: np.ndarray, alpha=1.0, gamma=0.0, size_average=True, ignore_index: int = 255, weight=None):
    """utils/loss.py:24-35 — per-pixel `ce = F.cross_entropy(reduction='none', ignore_index, weight)` (= w[y]*nll,
    0 where ignored), `pt = exp(-ce)`, `focal = alpha*(1-pt)**gamma*ce`, `.mean()` over ALL pixels or `.sum()`.
    Returns (loss float64, grad float64 like logits)."""
    x = np.asarray(logits, dtype=np.float64)
    y = np.asarray(labels).astype(np.int64)
    C = x.shape[1]
    xm = np.moveaxis(x, 1, -1).reshape(-1, C)
    yf = y.reshape(-1)
    w = np.ones(C, dtype=np.float64) if weight is None else np.asarray(weight, dtype=np.float64)
    valid = (yf != ignore_index) & (yf >= 0) & (yf < C)
    m = xm.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(xm - m).sum(axis=1))
    yc = np.where(valid, yf, 0)
    wy = np.where(valid, w[yc], 0.0)
    ce = wy * (lse - xm[np.arange(xm.shape[0]), yc])
    pt = np.exp(-ce)
    om = 1.0 - pt
    with np.errstate(invalid="ignore", divide="ignore"):
        mod = np.ones_like(om) if gamma == 0 else om ** gamma
        focal = alpha * mod * ce
        dmod = np.zeros_like(om) if gamma == 0 else np.where(om > 0, gamma * om ** (gamma - 1.0) * pt, 0.0)
    scale = 1.0 / yf.size if size_average else 1.0
    loss = focal.sum() * scale
    dfdce = alpha * (mod + ce * dmod)
    p = np.exp(xm - lse[:, None])
    onehot = np.zeros_like(p)
    onehot[np.arange(xm.shape[0]), yc] = 1.0
    g = (dfdce * wy)[:, None] * (p - onehot) * scale
    g[~valid] = 0.0
    g = np.moveaxis(g.reshape(x.shape[:1] + x.shape[2:] + (C,)), -1, 1)
    return float(loss), g


def to_tensor_normalize(img_u8_hwc: np.ndarray, mean, std) -> np.ndarray:
    """utils/ext_transforms.py:273-293 (ExtToTensor -> F.to_tensor: HWC uint8 -> CHW float32 / 255) followed by
    :298-324 (ExtNormalize -> F.normalize: (t - mean) / std), all in fp32 with IEEE division."""
    t = np.ascontiguousarray(np.asarray(img_u8_hwc).transpose(2, 0, 1)).astype(np.float32) / np.float32(255.0)
    mu = np.asarray(mean, dtype=np.float32)[:, None, None]
    sd = np.asarray(std, dtype=np.float32)[:, None, None]
    return ((t - mu) / sd).astype(np.float32)


def crop_flip(a: np.ndarray, x0: int, y0: int, H: int, W: int, flip: bool) -> np.ndarray:
    """utils/ext_transforms.py:366-393 (ExtRandomCrop -> F.crop(img, i, j, h, w)) then :94-111
    (ExtRandomHorizontalFlip -> F.hflip) on an HWC / HW array."""
    w = a[y0:y0 + H, x0:x0 + W]
    return w[:, ::-1] if flip else w


def adam_steps(p, grads, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
    """torch.optim.Adam / AdamW single-tensor rule (what train.py:432-441 constructs, torch defaults), in fp32 like
    torch: returns the parameter after applying `grads` (a sequence of gradient arrays) one step each."""
    f = np.float32
    p = np.asarray(p, dtype=f).copy()
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    b1, b2 = betas
    for t, g in enumerate(grads, 1):
        g = np.asarray(g, dtype=f)
        if decoupled:
            p = p * f(1.0 - lr * weight_decay)
        elif weight_decay != 0:
            g = g + f(weight_decay) * p
        m = m + (g - m) * f(1.0 - b1)
        v = v * f(b2) + f(1.0 - b2) * g * g
        bc1, bc2 = 1.0 - b1 ** t, 1.0 - b2 ** t
        denom = np.sqrt(v) / f(np.sqrt(bc2)) + f(eps)
        p = p - f(lr / bc1) * (m / denom)
    return p


def upsample_bilinear_nchw(x: np.ndarray, H: int, W: int) -> np.ndarray:
    """network/utils.py:22 — F.interpolate(x, size, mode='bilinear', align_corners=False) on NCHW fp32."""
    x = np.asarray(x, dtype=np.float32)
    Hi, Wi = x.shape[-2:]

    def src(o, n_in, n_out):
        s = (np.arange(n_out, dtype=np.float32) + np.float32(0.5)) * np.float32(n_in / n_out) - np.float32(0.5)
        s = np.maximum(s, np.float32(0))
        i0 = np.minimum(s.astype(np.int64), n_in - 1)
        i1 = i0 + (i0 < n_in - 1)
        return i0, i1, (s - i0.astype(np.float32)).astype(np.float32)

    y0, y1, ly = src(None, Hi, H)
    x0, x1, lx = src(None, Wi, W)
    ly = ly[:, None]
    lx = lx[None, :]
    a = x[..., y0, :][..., :, x0]
    b = x[..., y0, :][..., :, x1]
    c = x[..., y1, :][..., :, x0]
    d = x[..., y1, :][..., :, x1]
    return ((1 - ly) * (1 - lx) * a + (1 - ly) * lx * b + ly * (1 - lx) * c + ly * lx * d).astype(np.float32)
