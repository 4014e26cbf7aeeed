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


# ---------------------------------------------------------------------------------------------
# ExtRandomScale / ExtRandomCrop(pad_if_needed). The arithmetic lives in a third-party
# dependency: Pillow (requirements.txt pins 10.x; 12.2 here - src/libImaging/Resample.c and Geometry.c are unchanged
# between them for this path), reached through torchvision's F.resize -> PIL.Image.resize. Restated from the
# published algorithm and pinned against PIL itself in tests/test_scale_oracle.py (PIL is importable on the build
# container AND on the GPU box) plus the reference-generated fixtures tests/golden/scale_rows.npz.

_PIL_PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1.0) over the full box
    (0, in_size): returns (ksize, bounds int32 [out,2] = (xmin, count), kk int32 [out, ksize] fixed point 2^-22)."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = 0.0 + (xx + 0.5) * scale
    ss = 1.0 / filterscale
    xmin = (center - support + 0.5).astype(np.int64)            # C (int) cast: truncation toward zero
    xmin = np.maximum(xmin, 0)
    xmax = (center + support + 0.5).astype(np.int64)
    xmax = np.minimum(xmax, in_size) - xmin
    x = np.arange(ksize, dtype=np.int64)[None, :]
    arg = ((x + xmin[:, None]).astype(np.float64) - center[:, None] + 0.5) * ss
    arg = np.abs(arg)
    w = np.where(arg < 1.0, 1.0 - arg, 0.0)
    w = np.where(x < xmax[:, None], w, 0.0)
    # the C loop adds the weights one by one, left to right (np.sum would pair them)
    acc = np.zeros(out_size, dtype=np.float64)
    for k in range(ksize):
        acc = acc + w[:, k]
    ww = acc[:, None]
    kd = np.where(ww != 0.0, w / np.where(ww != 0.0, ww, 1.0), w)
    kd = np.where(x < xmax[:, None], kd, 0.0)
    kk = np.where(kd < 0, (-0.5 + kd * (1 << _PIL_PRECISION_BITS)), (0.5 + kd * (1 << _PIL_PRECISION_BITS))).astype(np.int64).astype(np.int32)
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return ksize, bounds, kk


def _pil_clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> _PIL_PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), BILINEAR) on a uint8 HWC (or HW) array: Resample.c ImagingResampleInner - the
    horizontal pass into a uint8 intermediate, then the vertical pass, each `clip8((1 << 21) + sum(px * k))`."""
    a = np.asarray(img)
    squeeze = a.ndim == 2
    if squeeze:
        a = a[:, :, None]
    H, W, _ = a.shape
    if (out_h, out_w) == (H, W):
        return np.asarray(img).copy()
    half = 1 << (_PIL_PRECISION_BITS - 1)
    cur = a
    if out_w != W:
        ks, b, kk = pil_bilinear_coeffs(W, out_w)
        idx = np.minimum(b[:, 0:1] + np.arange(ks)[None, :], W - 1)          # taps beyond the count carry k = 0
        g = cur[:, idx, :].astype(np.int64)                                   # [H, out_w, ks, C]
        cur = _pil_clip8(half + (g * kk[None, :, :, None].astype(np.int64)).sum(axis=2))
    if out_h != H:
        ks, b, kk = pil_bilinear_coeffs(H, out_h)
        idx = np.minimum(b[:, 0:1] + np.arange(ks)[None, :], H - 1)
        g = cur[idx, :, :].astype(np.int64)                                   # [out_h, ks, W', C]
        cur = _pil_clip8(half + (g * kk[:, :, None, None].astype(np.int64)).sum(axis=1))
    return cur[:, :, 0] if squeeze else cur


def pil_nearest_table(in_size: int, out_size: int) -> np.ndarray:
    """Geometry.c ImagingScaleAffine: xo = a*0.5, then `xin = (int) xo; xo += a` - the source index of every output
    pixel comes from a RUNNING double sum (not from (x + 0.5) * a), which decides exact-boundary cases."""
    a = float(np.float32(in_size) - np.float32(0.0)) / out_size
    steps = np.full(out_size, a, dtype=np.float64)
    steps[0] = 0.0 + a * 0.5
    xo = np.cumsum(steps)                                                     # sequential adds, like the C loop
    xin = np.where(xo < 0.0, -1, xo.astype(np.int64))
    return xin.astype(np.int32)


def pil_resize_nearest(lbl: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), NEAREST) on an HW array (labels, utils/ext_transforms.py:109)."""
    a = np.asarray(lbl)
    H, W = a.shape[:2]
    if (out_h, out_w) == (H, W):
        return a.copy()
    yt, xt = pil_nearest_table(H, out_h), pil_nearest_table(W, out_w)
    out = np.zeros((out_h, out_w) + a.shape[2:], dtype=a.dtype)
    oky, okx = (yt >= 0) & (yt < H), (xt >= 0) & (xt < W)
    out[np.ix_(oky, okx)] = a[np.ix_(yt[oky], xt[okx])]
    return out


def random_scale_geometry(Hs: int, Ws: int, scale: float, crop_hw, pad_if_needed: bool = True):
    """Sizes the reference's train pipeline goes through for one sample: ExtRandomScale (utils/ext_transforms.py:107:
    target = (int(h*scale), int(w*scale))), then ExtRandomCrop's pad_if_needed (:377-385: F.pad on ALL four sides by
    int((1 + tw - w) / 2) when too narrow, then by int((1 + th - h) / 2) when still too low, fill 0).
    Returns (sh, sw, pad, Hp, Wp): scaled size, total padding per side, padded size."""
    sh, sw = int(Hs * scale), int(Ws * scale)
    th, tw = crop_hw
    pad = 0
    if pad_if_needed and sw + 2 * pad < tw:
        pad += int((1 + tw - (sw + 2 * pad)) / 2)
    if pad_if_needed and sh + 2 * pad < th:
        pad += int((1 + th - (sh + 2 * pad)) / 2)
    return sh, sw, pad, sh + 2 * pad, sw + 2 * pad


def random_scale_crop(img_u8: np.ndarray, lbl_u8: np.ndarray, sh: int, sw: int, pad: int, y0: int, x0: int, H: int, W: int,
                      flip: bool, mean, std):
    """ExtRandomScale -> ExtRandomCrop(pad_if_needed) -> ExtRandomHorizontalFlip -> ExtToTensor -> ExtNormalize
    (train.py:355-362) for one sample with the random draws given: float32 [3,H,W] image, uint8 [H,W] label."""
    si = pil_resize_bilinear_u8(img_u8, sh, sw)
    sl = pil_resize_nearest(lbl_u8, sh, sw)
    if pad:
        si = np.pad(si, ((pad, pad), (pad, pad), (0, 0)))
        sl = np.pad(sl, ((pad, pad), (pad, pad)))
    ci, cl = crop_flip(si, x0, y0, H, W, flip), crop_flip(sl, x0, y0, H, W, flip)
    return to_tensor_normalize(ci, mean, std), np.ascontiguousarray(cl)
