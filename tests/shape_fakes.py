"""numpy stand-ins for the four shape-metric entry points of iswm_b200.ops, built on oracle/shape_np.py - TEST INFRASTRUCTURE.
The CPU tests run the host-side evaluators (iswm_b200/metrics/shape_metrics.py) on these in place of the CUDA kernels, so that the
host arithmetic is pinned to the reference fixtures without a GPU; the GPU tests hold the real kernels to the same functions."""
import numpy as np
import torch

from oracle import shape_np as S


def mask_preprocess(mask, min_valid_area=None):
    m = mask.cpu().numpy()
    if m.ndim == 2:
        m = m[None]
    N, H, W = m.shape
    thr = H * W * 0.001 if min_valid_area is None else min_valid_area
    support, front, info = np.zeros((N, H, W), np.uint8), np.full((N, H), -1, np.int32), np.zeros((N, 8), np.int32)
    for n in range(N):
        b = (m[n] > 0).astype(np.uint8)
        b = S.box_dilate(S.box_erode(S.box_erode(S.box_dilate(b, 1), 1), 1), 1)
        lab, areas, _ = S.label8(b)
        valid = np.where(areas >= thr)[0]
        info[n, 0], info[n, 1], info[n, 4] = areas.size, valid.size, -1
        if valid.size:
            best = valid[np.argmax(areas[valid])]
            sup = lab == best + 1
            support[n] = sup
            info[n, 2], info[n, 3] = areas[best], sup.sum()
            info[n, 4] = n * H * W + int(np.flatnonzero(sup)[0])
            for i in range(H):
                w = np.flatnonzero(sup[i])
                if w.size:
                    front[n, i] = w[0]
    return torch.from_numpy(support), torch.from_numpy(front), torch.from_numpy(info)


def region_components(pred, gt, min_area=50, cap=1024):
    p, g = pred.cpu().numpy(), gt.cpu().numpy()
    if p.ndim == 2:
        p, g = p[None], g[None]
    N = p.shape[0]
    counts, areas = np.zeros((N, 8), np.int32), np.zeros((N, cap), np.int32)
    for n in range(N):
        a, b = (p[n] > 0).astype(np.uint8), (g[n] > 0).astype(np.uint8)
        r = S.box_erode(S.box_dilate(a, 3), 2)
        _, ar, _ = S.label8(r)
        big = ar[ar >= min_area]
        counts[n, :6] = [a.sum(), b.sum(), (r & b).sum(), (r | b).sum(), big.size, ar.size]
        areas[n, :min(cap, big.size)] = big[:cap]
    return torch.from_numpy(counts), torch.from_numpy(areas)


def front_nearest(fa, fb):
    a, b = fa.cpu().numpy(), fb.cpu().numpy()
    N, H = a.shape
    d2, dx = np.full((N, H), -1, np.int32), np.full((N, H), -1, np.int32)
    for n in range(N):
        rows = np.flatnonzero(b[n] >= 0)
        for i in np.flatnonzero(a[n] >= 0):
            if rows.size:
                d = (i - rows).astype(np.int64) ** 2 + (a[n, i] - b[n, rows]).astype(np.int64) ** 2
                j = int(np.argmin(d))                      # first minimum, like the strict `<` of the reference loop
                d2[n, i], dx[n, i] = d[j], abs(int(a[n, i]) - int(b[n, rows[j]]))
    return torch.from_numpy(d2), torch.from_numpy(dx)


def front_window_diff(front, other, window):
    f, o = front.cpu().numpy(), other.cpu().numpy()
    N, H, W = o.shape
    out = np.full((N, H), -1, np.int32)
    for n in range(N):
        for i in np.flatnonzero(f[n] >= 0):
            cf = int(f[n, i])
            s, e = max(0, cf - window), min(W, cf + window)
            w = np.flatnonzero(o[n, i, s:e])
            if w.size:
                out[n, i] = abs(cf - (w[0] + s))
    return torch.from_numpy(out)


def install(monkeypatch):
    from iswm_b200 import ops
    from iswm_b200.metrics import shape_metrics
    for name in ("mask_preprocess", "region_components", "front_nearest", "front_window_diff"):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(shape_metrics.MaskUtils, "device", "cpu")
