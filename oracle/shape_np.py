"""CPU oracle for the per-image shape / front / temporal evaluators - TEST INFRASTRUCTURE ONLY.

A numpy restatement of metrics/utils/mask_utils.py, front_tracking_metrics.py, region_metrics.py and
temporal_metrics.py (reference paths), each function citing the lines it follows. Only tests/ may import it.

Third-party arithmetic restated here (the reference calls it, it is not in the reference tree):
  OpenCV (requirements.txt: opencv-python 4.x; 4.13 here) cv2.morphologyEx / dilate / erode with a 3x3 box and the
    default border (pixels outside the image are ignored) and cv2.connectedComponentsWithStats (8-connectivity; the
    LABEL ORDER of its block-based scan - components numbered by their first 2x2 block in block-raster order - matters
    because the reference breaks area ties with np.argmax);
  SciPy scipy.ndimage.label with the 8-connected structure (only region areas are used).
Pinned against the REAL reference classes run in the build container: tests/golden/shape_rows.npz
(oracle/gen_golden_shape.py) and, where cv2 is importable, a direct comparison in tests/test_shape_oracle.py.
The restatement itself needs numpy only (a small union-find labelling), so it also runs where cv2 / scipy are absent.
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------------------------------------ primitives
def box_dilate(m: np.ndarray, r: int) -> np.ndarray:
    """cv2.dilate with a (2r+1)^2 box, default border: max over the in-image neighbours."""
    H, W = m.shape
    p = np.zeros((H + 2 * r, W + 2 * r), dtype=m.dtype)
    p[r:r + H, r:r + W] = m
    out = np.zeros_like(m)
    for dy in range(2 * r + 1):
        for dx in range(2 * r + 1):
            out = np.maximum(out, p[dy:dy + H, dx:dx + W])
    return out


def box_erode(m: np.ndarray, r: int) -> np.ndarray:
    """cv2.erode with a (2r+1)^2 box, default border: min over the in-image neighbours."""
    H, W = m.shape
    p = np.ones((H + 2 * r, W + 2 * r), dtype=m.dtype)
    p[r:r + H, r:r + W] = m
    out = np.ones_like(m)
    for dy in range(2 * r + 1):
        for dx in range(2 * r + 1):
            out = np.minimum(out, p[dy:dy + H, dx:dx + W])
    return out


def label8(m: np.ndarray):
    """8-connected components of a binary mask: (labels int32 with 0 = background, areas [n], block keys [n]); labels
    are numbered 1..n in OpenCV's order: by the first 2x2 block of the component in block-raster order."""
    H, W = m.shape
    idx = np.arange(H * W, dtype=np.int64).reshape(H, W)
    parent = np.where(m > 0, idx, -1).reshape(-1)
    fg = m > 0
    # iterate "take the minimum label over the 3x3 neighbourhood" + pointer jumping until nothing changes
    lab = np.where(fg, idx, H * W).astype(np.int64)
    while True:
        p = np.full((H + 2, W + 2), H * W, dtype=np.int64)
        p[1:-1, 1:-1] = lab
        nb = lab.copy()
        for dy in range(3):
            for dx in range(3):
                nb = np.minimum(nb, p[dy:dy + H, dx:dx + W])
        nb = np.where(fg, nb, H * W)
        # pointer jumping through the flat table
        flat = nb.reshape(-1).copy()
        sel = flat < H * W
        for _ in range(32):
            nxt = flat.copy()
            nxt[sel] = flat[flat[sel]]
            if np.array_equal(nxt, flat):
                break
            flat = nxt
        new = flat.reshape(H, W)
        if np.array_equal(new, lab):
            break
        lab = new
    del parent
    roots = np.unique(lab[fg])
    if roots.size == 0:
        return np.zeros((H, W), np.int32), np.zeros(0, np.int64), np.zeros(0, np.int64)
    yy, xx = np.nonzero(fg)
    key = (yy // 2) * ((W + 1) // 2) + (xx // 2)
    comp = np.searchsorted(roots, lab[fg])
    areas = np.bincount(comp, minlength=roots.size)
    keys = np.full(roots.size, np.iinfo(np.int64).max)
    np.minimum.at(keys, comp, key)
    order = np.argsort(keys, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    out = np.zeros((H, W), np.int32)
    out[fg] = rank[comp] + 1
    return out, areas[order], keys[order]


# ------------------------------------------------------------------------------------------------ mask_utils.py
def preprocess_mask(mask):
    """metrics/utils/mask_utils.py:7-52. Returns the array the reference returns: uint8 0/1, or float64 0/weight when more
    than one valid region exists."""
    mask = np.asarray(mask)
    if mask.ndim == 3:
        mask = mask[-1]
    m = (mask > 0).astype(np.uint8)
    m = box_erode(box_dilate(m, 1), 1)                       # MORPH_CLOSE
    m = box_dilate(box_erode(m, 1), 1)                       # MORPH_OPEN
    labels, areas, _ = label8(m)
    if areas.size >= 1:                                       # num_labels > 1
        min_valid_area = m.size * 0.001
        valid_labels = np.where(areas >= min_valid_area)[0] + 1
        if len(valid_labels) > 0:
            largest = valid_labels[np.argmax(areas[valid_labels - 1])]
            base = (labels == largest).astype(np.uint8)
            if len(valid_labels) > 1:
                weight = max(0.4, 1.0 - 0.2 * (len(valid_labels) - 1))
                return base * weight
            return base
        return np.zeros_like(m)
    return m


def find_front_positions(mask):
    """mask_utils.py:54-76: (row, leftmost column) of every row holding a pixel EQUAL TO 1 of the re-preprocessed mask."""
    m = preprocess_mask(mask)
    if not np.any(m):
        return []
    out = []
    for i in range(m.shape[0]):
        w = np.where(m[i] == 1)[0]
        if len(w) > 0:
            out.append((i, w[0]))
    return out


def calculate_motion(curr_pred, prev_pred):
    """mask_utils.py:78-103."""
    cf, pf = find_front_positions(curr_pred), find_front_positions(prev_pred)
    if not cf or not pf:
        return 0.0
    cy, cx = np.mean([y for y, x in cf]), np.mean([x for y, x in cf])
    py, px = np.mean([y for y, x in pf]), np.mean([x for y, x in pf])
    distance = np.sqrt((cy - py) ** 2 + (cx - px) ** 2)
    return 1.0 / (1.0 + distance / (curr_pred.shape[0] * 0.1))


def calculate_stability(curr_pred, prev_pred):
    """mask_utils.py:105-135."""
    c, p = preprocess_mask(curr_pred), preprocess_mask(prev_pred)
    window = int(c.shape[1] * 0.1)
    scores = []
    for i in range(c.shape[0]):
        cp = np.where(c[i] == 1)[0]
        if len(cp) > 0:
            cf = cp[0]
            s, e = max(0, cf - window), min(c.shape[1], cf + window)
            pp = np.where(p[i, s:e] == 1)[0]
            if len(pp) > 0:
                scores.append(1.0 / (1.0 + abs(cf - (pp[0] + s)) / window))
    return np.mean(scores) if scores else 0.0


def check_wave_presence(mask, threshold=0.005):
    """mask_utils.py:137-142."""
    m = preprocess_mask(mask)
    return np.sum(m) / m.size >= threshold


# ------------------------------------------------------------------------------------------------ front_tracking_metrics.py
def front_tracking_error(pred, gt, max_distance_threshold=None):
    """front_tracking_metrics.py:17-109 (calculate_error)."""
    tau = pred.shape[1] * 0.1 if max_distance_threshold is None else max_distance_threshold
    pred, gt = preprocess_mask(pred), preprocess_mask(gt)
    pf, gf = find_front_positions(pred), find_front_positions(gt)
    if gf and not pf:
        return tau * 2.0
    if not gf and pf:
        return tau * 1.5
    if not gf and not pf:
        return 0.0

    def one_way(a, b):
        err = wsum = 0
        valid = 0
        for ay, ax in a:
            md, mdx = float("inf"), float("inf")
            for by, bx in b:
                d = np.sqrt((ay - by) ** 2 + (ax - bx) ** 2)
                if d < md:
                    md, mdx = d, abs(ax - bx)
            if md < tau:
                w = 1.0 / (mdx + 1e-6)
                err += md * w
                wsum += w
                valid += 1
        return err, wsum, valid

    pe, pw, pv = one_way(pf, gf)
    ge, gw, gv = one_way(gf, pf)
    if pv == 0 or gv == 0:
        return tau * 2.0
    pavg = pe / pw if pw > 0 else float("inf")
    gavg = ge / gw if gw > 0 else float("inf")
    coverage = gv / len(gf)
    return max(pavg, gavg) + (1.0 - coverage) * tau * 0.5


# ------------------------------------------------------------------------------------------------ region_metrics.py
def region_metrics(pred, gt, min_area_threshold=50):
    """region_metrics.py:6-115 (repair_small_gaps, _calculate_shape_metrics, _calculate_fragmentation_score,
    calculate_region_metrics). None = the reference's "invalid case"."""
    pred = (np.asarray(pred) > 0).astype(np.uint8)
    gt = (np.asarray(gt) > 0).astype(np.uint8)
    if np.sum(pred) == 0 or np.sum(gt) == 0:
        return None
    pred = box_erode(box_dilate(pred, 3), 2)                 # dilate x3, erode x2 with a 3x3 box
    inter, union = np.logical_and(pred, gt).sum(), np.logical_or(pred, gt).sum()
    similarity = inter / union
    _, areas, _ = label8(pred)
    regions = [int(a) for a in areas if a >= min_area_threshold]
    if not regions:
        frag, n = 0.0, 0
    else:
        srt = sorted(regions, reverse=True)
        total = sum(regions)
        ratios = [a / total for a in srt]
        frag = ratios[0]
        if len(regions) > 1:
            frag -= sum(r * (i + 1) / len(regions) for i, r in enumerate(ratios[1:])) * 0.5
        frag, n = float(max(0.0, min(1.0, frag))), len(regions)
    return {"fragmentation_score": frag, "similarity_score": float(similarity), "num_regions": n,
            "final_score": float(0.7 * frag + 0.3 * float(similarity))}


# ------------------------------------------------------------------------------------------------ temporal_metrics.py
class TemporalOracle:
    """temporal_metrics.py:5-181 (update / _calculate_sequence_temporal_consistency and its three branches)."""

    def __init__(self, sequence_length=7, threshold=0.005):
        self.sequence_length, self.threshold = sequence_length, threshold
        self.reset()

    def reset(self):
        self.preds, self.gts, self.temporal_scores = [], [], []
        self.transition_scores, self.stability_scores, self.motion_scores, self.wave_segment_scores = [], [], [], []

    def _transitions(self, gt_has, pred_has):
        gt_t, pr_t = np.diff(gt_has).astype(int), np.diff(pred_has).astype(int)
        if not np.any(gt_t):
            s = 1.0 if not np.any(pr_t) else 0.0
        else:
            gi, pi = np.where(gt_t)[0], np.where(pr_t)[0]
            s = 0.0 if len(pi) != len(gi) else 1.0 / (1.0 + np.mean(np.abs(gi - pi)))
        self.transition_scores.append(s)
        return s

    def _sequence(self, P, G):
        gt_has = [check_wave_presence(f, self.threshold) for f in G]
        pred_has = [check_wave_presence(f, self.threshold) for f in P]
        if not any(gt_has):
            return 1.0 - sum(pred_has) / len(pred_has)
        if all(gt_has):
            st = [calculate_stability(P[t], P[t - 1]) for t in range(1, len(P))]
            mo = [calculate_motion(P[t], P[t - 1]) for t in range(1, len(P))]
            self.stability_scores.append(np.mean(st) if st else 0.0)
            self.motion_scores.append(np.mean(mo) if mo else 0.0)
            return np.mean([0.5 * s + 0.5 * m for s, m in zip(st, mo)]) if st else 0.0
        tr = self._transitions(gt_has, pred_has)
        ws = [0.5 * calculate_stability(P[t], P[t - 1]) + 0.5 * calculate_stability(P[t], G[t]) for t in range(1, len(P)) if gt_has[t]]
        seg = np.mean(ws) if ws else 0.0
        self.wave_segment_scores.append(seg)
        return 0.6 * tr + 0.4 * seg

    def update(self, pred, gt):
        if pred.ndim > 2:
            pred = preprocess_mask(pred)
        if gt.ndim > 2:
            gt = preprocess_mask(gt)
        self.preds.append(pred)
        self.gts.append(gt)
        score = None
        if len(self.preds) == self.sequence_length:
            score = self._sequence(self.preds, self.gts)
            self.temporal_scores.append(score)
            self.preds, self.gts = self.preds[1:], self.gts[1:]
        return score

    def get_mean_score(self):
        return np.mean(self.temporal_scores) if self.temporal_scores else 0.0

    def get_detailed_statistics(self):
        m = lambda a: np.mean(a) if a else 0.0
        return {"mean_score": self.get_mean_score(), "mean_transition": m(self.transition_scores), "mean_stability": m(self.stability_scores),
                "mean_motion": m(self.motion_scores), "mean_wave_segment": m(self.wave_segment_scores), "score_count": len(self.temporal_scores)}
