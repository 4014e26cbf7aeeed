"""Device-backed evaluators with the reference's interfaces:

    MaskUtils              metrics/utils/mask_utils.py:6-142
    FrontTrackingMetrics   metrics/front_tracking_metrics.py:6-133
    RegionMetrics          metrics/region_metrics.py:14-157
    TemporalMetrics        metrics/temporal_metrics.py:5-181

The image processing (morphology, connected components, front scans, nearest-point and window searches - cv2 / scipy /
python double loops in the reference) runs in libiswm_b200's kernels (csrc/shape_metrics.cu), integer and bit-exact; the
float64 arithmetic on their results is done here on the host with the reference's own numpy calls in the reference's
operation order, so every score equals the reference's to the last bit (tests/test_shape_gpu.py against fixtures produced
by the real classes). There is no CPU path: masks that are not CUDA tensors are copied to the device first.

A preprocessed mask is kept as `Pre` (support on the device + weight) rather than the reference's `support * weight` array:
`mask > 0` is the support, `mask == 1` is the support when weight == 1 and nothing otherwise (the reference's weighted
masks - several valid regions - hold 0.8 / 0.6 / 0.4, which its `== 1` tests never match), sum(mask) = count * weight.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from .. import ops


def _dev(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("the shape / front / temporal evaluators need a CUDA device (iswm_b200 has no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_mask(a, device) -> torch.Tensor:
    """[H,W] or [T,H,W] integer mask on the device (uint8 / int32 / int64 kept as they are)."""
    if isinstance(a, Pre):
        return a.support
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.ascontiguousarray(a)
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        elif a.dtype.kind == "f":
            a = (a > 0).astype(np.uint8)
        elif a.dtype not in (np.uint8, np.int32, np.int64):
            a = a.astype(np.int64)
        t = torch.from_numpy(a)
    if t.is_floating_point() or t.dtype == torch.bool:
        t = (t > 0).to(torch.uint8)
    elif t.dtype not in (torch.uint8, torch.int32, torch.int64):
        t = t.long()
    return t.to(device, non_blocking=True)


@dataclass
class Pre:
    """MaskUtils.preprocess_mask's result: the array the reference returns is `support * weight`."""
    support: torch.Tensor            # uint8 [H,W] on the device
    front: np.ndarray                # int32 [H] host: leftmost support column per row, -1 = none
    components: int
    valid: int
    count: int                       # support pixels

    @property
    def shape(self):
        return tuple(self.support.shape)

    @property
    def weight(self) -> float:
        return 1.0 if self.valid <= 1 else max(0.4, 1.0 - 0.2 * (self.valid - 1))

    @property
    def ones_front(self) -> np.ndarray:
        """Fronts of the pixels EQUAL TO 1 (what every `== 1` scan of the reference sees)."""
        return self.front if self.valid <= 1 else np.full_like(self.front, -1)

    def front_positions(self):
        f = self.ones_front
        rows = np.nonzero(f >= 0)[0]
        return [(int(i), np.int64(f[i])) for i in rows]

    def numpy(self) -> np.ndarray:
        """The reference's return value: uint8 0/1, or float64 0/weight when several valid regions exist (mask_utils.py:38-43)."""
        s = self.support.cpu().numpy()
        return s if self.valid <= 1 else s * self.weight


class MaskUtils:
    """metrics/utils/mask_utils.py:6-142; `device` is a class attribute (None = the current CUDA device)."""

    device = None

    @staticmethod
    def preprocess(mask) -> Pre:
        return MaskUtils.preprocess_many([mask])[0]

    @staticmethod
    def preprocess_many(masks) -> List[Pre]:
        """One launch sequence for several frames of one size ([T,H,W] inputs contribute their LAST frame, mask_utils.py:12-13)."""
        dev = _dev(MaskUtils.device)
        ts = []
        for m in masks:
            t = _to_mask(m, dev)
            ts.append(t[-1] if t.dim() == 3 else t)
        dt = ts[0].dtype if all(t.dtype == ts[0].dtype for t in ts) else torch.int64
        batch = torch.stack([t.to(dt) for t in ts])
        support, front, info = ops.mask_preprocess(batch)
        info_h, front_h = info.cpu().numpy(), front.cpu().numpy()
        return [Pre(support[i], front_h[i], int(info_h[i, 0]), int(info_h[i, 1]), int(info_h[i, 3])) for i in range(len(ts))]

    @staticmethod
    def preprocess_mask(mask) -> np.ndarray:
        """Drop-in: returns the numpy array the reference returns."""
        return MaskUtils.preprocess(mask).numpy()

    @staticmethod
    def find_front_positions(mask):
        """mask_utils.py:54-76 (the mask is preprocessed again, as there)."""
        return MaskUtils.preprocess(mask).front_positions()

    @staticmethod
    def _motion(c: Pre, p: Pre, height: int) -> float:
        cf, pf = c.front_positions(), p.front_positions()
        if not cf or not pf:
            return 0.0
        curr_y, curr_x = np.mean([y for y, x in cf]), np.mean([x for y, x in cf])
        prev_y, prev_x = np.mean([y for y, x in pf]), np.mean([x for y, x in pf])
        distance = np.sqrt((curr_y - prev_y) ** 2 + (curr_x - prev_x) ** 2)
        return 1.0 / (1.0 + distance / (height * 0.1))

    @staticmethod
    def calculate_motion(curr_pred, prev_pred):
        """mask_utils.py:78-103."""
        c, p = MaskUtils.preprocess_many([curr_pred, prev_pred])
        return MaskUtils._motion(c, p, c.shape[0])

    @staticmethod
    def _stability_many(pairs) -> List[float]:
        """pairs of (Pre curr, Pre prev), all of one size: one window-search launch for all of them."""
        if not pairs:
            return []
        H, W = pairs[0][0].shape
        window = int(W * 0.1)
        dev = pairs[0][0].support.device
        fronts = torch.from_numpy(np.stack([c.ones_front for c, _ in pairs])).to(dev)
        zero = None
        others = []
        for _, p in pairs:
            if p.valid <= 1:
                others.append(p.support)
            else:                                             # a weighted mask holds no pixel equal to 1
                zero = torch.zeros_like(p.support) if zero is None else zero
                others.append(zero)
        diff = ops.front_window_diff(fronts, torch.stack(others), window).cpu().numpy()
        out = []
        for k in range(len(pairs)):
            d = diff[k][diff[k] >= 0].astype(np.int64)
            scores = [1.0 / (1.0 + v / window) for v in d]
            out.append(np.mean(scores) if scores else 0.0)
        return out

    @staticmethod
    def calculate_stability(curr_pred, prev_pred):
        """mask_utils.py:105-135."""
        c, p = MaskUtils.preprocess_many([curr_pred, prev_pred])
        return MaskUtils._stability_many([(c, p)])[0]

    @staticmethod
    def _presence(p: Pre, threshold: float) -> bool:
        """np.sum(mask) / mask.size >= threshold (mask_utils.py:139-142). For an unweighted mask the sum is the integer pixel count; a
        weighted mask (several valid regions) is summed by numpy in its pairwise order, which `count * weight` reproduces to ~1e-13
        relative - when the ratio sits that close to the threshold the mask itself is fetched and summed exactly as the reference does."""
        H, W = p.shape
        if p.valid <= 1:
            return bool(p.count / (H * W) >= threshold)
        ratio = p.count * p.weight / (H * W)
        if abs(ratio - threshold) <= 1e-9 * max(abs(threshold), 1e-30):
            m = p.numpy()
            ratio = np.sum(m) / m.size
        return bool(ratio >= threshold)

    @staticmethod
    def check_wave_presence(mask, threshold=0.005):
        """mask_utils.py:137-142."""
        return MaskUtils._presence(MaskUtils.preprocess(mask), threshold)


class FrontTrackingMetrics:
    """metrics/front_tracking_metrics.py:6-133."""

    accepts_device = True

    def __init__(self):
        self.max_distance_threshold = None
        self.tracking_errors = []

    def set_max_distance_threshold(self, image_width):
        self.max_distance_threshold = image_width * 0.1

    @staticmethod
    def _one_way(d2: np.ndarray, dx: np.ndarray, tau: float):
        """front_tracking_metrics.py:44-66: weighted error over the points closer than tau, accumulated in row order."""
        err = wsum = 0
        valid = 0
        for i in np.nonzero(d2 >= 0)[0]:
            min_dist = np.sqrt(np.int64(d2[i]))
            if min_dist < tau:
                weight = 1.0 / (np.int64(dx[i]) + 1e-6)
                err += min_dist * weight
                wsum += weight
                valid += 1
        return err, wsum, valid

    def calculate_error(self, pred, gt):
        """front_tracking_metrics.py:17-109: both masks are preprocessed, and find_front_positions preprocesses the result again."""
        shape = pred.shape
        if self.max_distance_threshold is None:
            self.set_max_distance_threshold(shape[1])
        p1, g1 = MaskUtils.preprocess_many([pred, gt])
        p2, g2 = MaskUtils.preprocess_many([p1, g1])
        return self._error_from(p2, g2)

    def _error_from(self, p2: Pre, g2: Pre):
        tau = self.max_distance_threshold
        pf, gf = p2.ones_front, g2.ones_front
        has_p, has_g = bool((pf >= 0).any()), bool((gf >= 0).any())
        if has_g and not has_p:
            return tau * 2.0
        if not has_g and has_p:
            return tau * 1.5
        if not has_g and not has_p:
            return 0.0
        dev = p2.support.device
        a = torch.from_numpy(np.stack([pf, gf])).to(dev)
        b = torch.from_numpy(np.stack([gf, pf])).to(dev)
        d2, dx = ops.front_nearest(a, b)
        d2, dx = d2.cpu().numpy(), dx.cpu().numpy()
        pe, pw, pv = self._one_way(d2[0], dx[0], tau)
        ge, gw, gv = self._one_way(d2[1], dx[1], tau)
        if pv == 0 or gv == 0:
            return tau * 2.0
        pred_avg = pe / pw if pw > 0 else float("inf")
        gt_avg = ge / gw if gw > 0 else float("inf")
        coverage = gv / int((gf >= 0).sum())
        return max(pred_avg, gt_avg) + (1.0 - coverage) * tau * 0.5

    def update(self, pred, gt):
        if self.max_distance_threshold is None:
            self.set_max_distance_threshold(pred.shape[1])
        error = self.calculate_error(pred, gt)
        if error is not None:
            self.tracking_errors.append(error)
        return error

    def get_mean_error(self):
        valid = [x for x in self.tracking_errors if x is not None and not np.isinf(x)]
        if not valid:
            if self.max_distance_threshold is not None:
                return self.max_distance_threshold * 2.0
            return float("inf")
        return np.mean(valid)

    def reset(self):
        self.tracking_errors = []


class RegionMetrics:
    """metrics/region_metrics.py:14-157."""

    accepts_device = True

    def __init__(self, device=None):
        self.valid_scores = []
        self.total_cases = 0
        self.invalid_cases = 0
        self.min_area_threshold = 50
        self.device = device

    def _calculate_fragmentation_score(self, regions):
        """region_metrics.py:21-37; regions = list of areas."""
        if not regions:
            return 0.0
        srt = sorted(regions, reverse=True)
        total = sum(regions)
        ratios = [a / total for a in srt]
        score = ratios[0]
        if len(regions) > 1:
            score -= sum(r * (i + 1) / len(regions) for i, r in enumerate(ratios[1:])) * 0.5
        return max(0.0, min(1.0, score))

    def calculate_region_metrics(self, pred, gt):
        """region_metrics.py:63-115."""
        dev = _dev(self.device if self.device is not None else MaskUtils.device)
        p, g = _to_mask(pred, dev), _to_mask(gt, dev)
        H, W = p.shape[-2:]
        cap = max(16, (H * W) // max(1, self.min_area_threshold) + 1)
        counts, areas = ops.region_components(p, g, self.min_area_threshold, cap)
        c = counts.cpu().numpy()[0]
        if c[0] == 0 or c[1] == 0:
            return None
        similarity = np.int64(c[2]) / np.int64(c[3])
        regions = [int(a) for a in areas[0, :min(int(c[4]), cap)].cpu().numpy()]
        if regions:
            frag, n = float(self._calculate_fragmentation_score(regions)), len(regions)
        else:
            frag, n = 0.0, 0
        return {"fragmentation_score": frag, "similarity_score": float(similarity), "num_regions": n,
                "final_score": float(0.7 * frag + 0.3 * float(similarity))}

    def update(self, pred, gt):
        self.total_cases += 1
        m = self.calculate_region_metrics(pred, gt)
        if m is not None:
            self.valid_scores.append(m["final_score"])
        else:
            self.invalid_cases += 1
        return m

    def get_mean_score(self):
        return np.mean(self.valid_scores) if self.valid_scores else 0.0

    def get_statistics(self):
        n = len(self.valid_scores)
        return {"mean_score": np.mean(self.valid_scores) if n else None, "total_cases": self.total_cases, "valid_cases": n,
                "invalid_cases": self.invalid_cases, "valid_ratio": n / self.total_cases if n else 0.0}

    def reset(self):
        self.valid_scores = []
        self.total_cases = 0
        self.invalid_cases = 0


class TemporalMetrics:
    """metrics/temporal_metrics.py:5-181. The reference re-preprocesses every queued frame inside every helper it calls on every
    window; here a frame is preprocessed when it is queued (the queue holds `Pre` objects of the frames AS THE HELPERS SEE THEM)
    and every window is evaluated from those - the same values, computed once."""

    accepts_device = True

    def __init__(self, sequence_length=7, threshold=0.005):
        self.sequence_length = sequence_length
        self.threshold = threshold
        self.reset()

    def reset(self):
        self.sequence_predictions: List[Pre] = []
        self.sequence_groundtruth: List[Pre] = []
        self.temporal_scores = []
        self.transition_scores = []
        self.stability_scores = []
        self.motion_scores = []
        self.wave_segment_scores = []

    def _evaluate_transitions(self, gt_has_wave, pred_has_wave):
        """temporal_metrics.py:21-43."""
        gt_t, pr_t = np.diff(gt_has_wave).astype(int), np.diff(pred_has_wave).astype(int)
        if not np.any(gt_t):
            score = 1.0 if not np.any(pr_t) else 0.0
        else:
            gi, pi = np.where(gt_t)[0], np.where(pr_t)[0]
            score = 0.0 if len(pi) != len(gi) else 1.0 / (1.0 + np.mean(np.abs(gi - pi)))
        self.transition_scores.append(score)
        return score

    def _calculate_sequence_temporal_consistency(self, P: List[Pre], G: List[Pre]):
        """temporal_metrics.py:110-126 and the three branches :45-108."""
        gt_has = [MaskUtils._presence(f, self.threshold) for f in G]
        pred_has = [MaskUtils._presence(f, self.threshold) for f in P]
        if not any(gt_has):
            return 1.0 - sum(pred_has) / len(pred_has)
        H = P[0].shape[0]
        if all(gt_has):
            st = MaskUtils._stability_many([(P[t], P[t - 1]) for t in range(1, len(P))])
            mo = [MaskUtils._motion(P[t], P[t - 1], H) for t in range(1, len(P))]
            self.stability_scores.append(np.mean(st) if st else 0.0)
            self.motion_scores.append(np.mean(mo) if mo else 0.0)
            return np.mean([0.5 * s + 0.5 * m for s, m in zip(st, mo)]) if st else 0.0
        tr = self._evaluate_transitions(gt_has, pred_has)
        ts = [t for t in range(1, len(P)) if gt_has[t]]
        st = MaskUtils._stability_many([(P[t], P[t - 1]) for t in ts] + [(P[t], G[t]) for t in ts])
        ws = [0.5 * st[k] + 0.5 * st[len(ts) + k] for k in range(len(ts))]
        seg = np.mean(ws) if ws else 0.0
        self.wave_segment_scores.append(seg)
        return 0.6 * tr + 0.4 * seg

    def update(self, pred, gt):
        """temporal_metrics.py:128-151: a 3-D input is replaced by its preprocessed LAST frame before it is queued, and every
        helper preprocesses the queued array again - the queue holds that second result (2-D inputs are queued raw: one pass)."""
        firsts = MaskUtils.preprocess_many([pred, gt])
        seconds = MaskUtils.preprocess_many(firsts)
        p = seconds[0] if len(pred.shape) > 2 else firsts[0]
        g = seconds[1] if len(gt.shape) > 2 else firsts[1]
        return self._push(p, g)

    def _push(self, p: Pre, g: Pre):
        self.sequence_predictions.append(p)
        self.sequence_groundtruth.append(g)
        score = None
        if len(self.sequence_predictions) == self.sequence_length:
            score = self._calculate_sequence_temporal_consistency(self.sequence_predictions, self.sequence_groundtruth)
            self.temporal_scores.append(score)
            self.sequence_predictions = self.sequence_predictions[1:]
            self.sequence_groundtruth = self.sequence_groundtruth[1:]
        return score

    def get_latest_score(self):
        return self.temporal_scores[-1] if self.temporal_scores else 0.0

    def get_mean_score(self):
        return np.mean(self.temporal_scores) if self.temporal_scores else 0.0

    def get_detailed_statistics(self):
        m = lambda a: np.mean(a) if a else 0.0   # noqa: E731
        return {"mean_score": self.get_mean_score(), "mean_transition": m(self.transition_scores), "mean_stability": m(self.stability_scores),
                "mean_motion": m(self.motion_scores), "mean_wave_segment": m(self.wave_segment_scores), "score_count": len(self.temporal_scores)}
