"""StreamMetrics — drop-in for metrics/stream_metrics.py:7-195 on the confusion-matrix path.

`_fast_hist`, the confusion-matrix accumulation and the IoU / precision / recall / F1 / MIoU
arithmetic follow the reference exactly (integer counts bit-exact, ratios in float64 with
eps=1e-7); counts are produced by the CUDA confusion kernel and kept ON DEVICE as int64 until
a result is read. The per-image shape / temporal / front-tracking evaluators of the reference
(metrics/region_metrics.py, temporal_metrics.py, front_tracking_metrics.py — cv2/scipy CPU
heuristics) are the device-backed classes of `iswm_b200.metrics.shape_metrics`
(bit-identical scores, tests/test_shape_gpu.py), created by default as stream_metrics.py:16-22 does.
They are also PLUG-IN slots: any object with the reference's evaluator interface (`update(pred, gt)`,
`reset()`, `get_mean_score()` / `get_mean_error()`) can be assigned to `temporal_evaluator`,
`region_evaluator`, `front_tracking_evaluator` (the reference's own classes work unchanged) and
`update()` feeds them exactly as stream_metrics.py:104-118 does. With `shape_metrics=False` the
slots start empty: an empty slot's result key is NaN ("not measured") and the weighted score is taken
over the measured terms with the reference's weights renormalised - a constant 0.0 would read as a
perfect Front Tracking Error and add +0.25 to every score.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import ops


def _to_host(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def _to_device_labels(a, device):
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.ascontiguousarray(a)
        if a.dtype not in (np.uint8, np.int32, np.int64):
            a = a.astype(np.int64)
        t = torch.from_numpy(a)
    if t.dtype not in (torch.uint8, torch.int32, torch.int64):
        t = t.long()
    return t.to(device, non_blocking=True).contiguous()


class StreamMetrics:
    def __init__(self, n_classes, sequence_length=7, temporal_stride=1, threshold=0.005, device=None, shape_metrics=True):
        self.n_classes = n_classes
        self.FOREGROUND_CLASS = 1
        self.sequence_length, self.temporal_stride, self.threshold = sequence_length, temporal_stride, threshold
        self.best_score = {"weighted_score": 0.0}
        # evaluators with the reference's interface (stream_metrics.py:16-22); None = not measured
        self.temporal_evaluator = None
        self.region_evaluator = None
        self.front_tracking_evaluator = None
        if shape_metrics:
            from .shape_metrics import FrontTrackingMetrics, RegionMetrics, TemporalMetrics
            self.temporal_evaluator = TemporalMetrics(sequence_length=sequence_length, threshold=threshold)
            self.region_evaluator = RegionMetrics()
            self.front_tracking_evaluator = FrontTrackingMetrics()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        self._cm_dev: Optional[torch.Tensor] = None
        self.process_group = None          # set by iswm_b200.parallel to all-reduce counts in get_results
        self.verbose = False               # the reference prints 6 lines per call (stream_metrics.py:41-46)

    # ---- device-side accumulation -------------------------------------------------------
    def _cm(self) -> torch.Tensor:
        if self.device is None:
            raise RuntimeError("StreamMetrics needs a CUDA device (no CPU fallback)")
        if self._cm_dev is None:
            self._cm_dev = torch.zeros(self.n_classes ** 2 + 1, dtype=torch.int64, device=self.device)
        return self._cm_dev

    def _fast_hist(self, label_true, label_pred):
        """stream_metrics.py:24-31 — returns the n x n int64 histogram of ONE call (host numpy)."""
        t = _to_device_labels(label_true, self.device).reshape(-1)
        p = _to_device_labels(label_pred, self.device).reshape(-1)
        cm = ops.confusion(t, p, self.n_classes)
        return cm[:-1].view(self.n_classes, self.n_classes).cpu().numpy()

    def update_cuda(self, label_trues: torch.Tensor, preds_or_logits: torch.Tensor, threshold: Optional[float] = None):
        """Tensor fast path: integer predictions [*] or logits [B,C,H,W] (argmax, or softmax[:,1] > threshold);
        counts stay on the device, nothing synchronises."""
        cm = self._cm()
        if preds_or_logits.is_floating_point():
            mode = 0 if threshold is None else 1
            ops.argmax_confusion(preds_or_logits, label_trues, mode=mode, threshold=0.5 if threshold is None else threshold, out=cm)
        else:
            ops.confusion(label_trues.reshape(-1), preds_or_logits.reshape(-1), self.n_classes, out=cm)

    def update(self, label_trues, label_preds, sequence_data=True):
        """stream_metrics.py:102-138: with sequence_data only the LAST frame of the window reaches the
        confusion matrix (:113-114); like the reference, every call re-evaluates the running results and
        refreshes `best_score` (:124-137) - one device->host read of the n x n counters per call, which is what the
        reference's numpy accumulator costs too. `update_cuda` is the non-synchronising path."""
        if sequence_data:
            if self.temporal_evaluator is not None:
                self.temporal_evaluator.update(label_preds, label_trues)
            t, p = label_trues[-1], label_preds[-1]
        else:
            t, p = label_trues, label_preds
        for ev in (self.region_evaluator, self.front_tracking_evaluator):
            if ev is not None:
                if getattr(ev, "accepts_device", False):
                    ev.update(p, t)
                else:
                    ev.update(_to_host(p), _to_host(t))
        t = _to_device_labels(t, self.device).reshape(-1)
        p = _to_device_labels(p, self.device).reshape(-1)
        ops.confusion(t, p, self.n_classes, out=self._cm())
        current = self.get_results(update_best=False)
        score = self._calculate_weighted_score(current)
        if score > self.best_score["weighted_score"]:
            self.best_score["weighted_score"] = score
            self.best_score.update({
                "miou": current["MIoU"], "foreground_iou": current["Foreground IoU"], "foreground_f1": current["Foreground F1"],
                "temporal_consistency": current["Temporal Consistency"], "front_tracking_error": current["Front Tracking Error"],
                "region_continuity": current["Region Continuity"],
            })

    @property
    def confusion_matrix(self) -> np.ndarray:
        """float64 [n,n] like the reference's np.zeros((n,n)) accumulator (read by train.py:739)."""
        cm = self._cm().clone()
        if self.process_group is not None:
            torch.distributed.all_reduce(cm, group=self.process_group)
        return cm[:-1].view(self.n_classes, self.n_classes).cpu().numpy().astype(np.float64)

    # ---- derived numbers ------------------------------------------------------------------
    def _calculate_foreground_metrics(self, hist):
        """stream_metrics.py:33-63."""
        fg = self.FOREGROUND_CLASS
        tp = hist[fg, fg]
        fp = hist[:, fg].sum() - tp
        fn = hist[fg, :].sum() - tp
        tn = hist.sum() - (tp + fp + fn)
        if self.verbose:
            print("\nConfusion Matrix Components:")
            print(f"True Positives: {tp}")
            print(f"False Positives: {fp}")
            print(f"False Negatives: {fn}")
            print(f"True Negatives: {tn}")
            print(f"Total Pixels: {hist.sum()}")
        eps = 1e-7
        foreground_iou = tp / (tp + fp + fn + eps)
        precision = tp / (tp + fp + eps)
        recall = tp / (tp + fn + eps)
        f1_score = 2 * precision * recall / (precision + recall + eps)
        btp = hist[0, 0]
        bfp = hist[:, 0].sum() - btp
        bfn = hist[0, :].sum() - btp
        background_iou = btp / (btp + bfp + bfn + eps)
        miou = (background_iou + foreground_iou) / 2.0
        return miou, foreground_iou, precision, recall, f1_score

    def _calculate_weighted_score(self, results):
        """stream_metrics.py:65-100 (weights 0.05 / 0.25 / 0.25 / 0.25 / 0.10 / 0.10); terms whose evaluator is not
        plugged in (NaN) are left out and the remaining weights renormalised."""
        fte = results["Front Tracking Error"]
        terms = [(0.05, results["MIoU"]), (0.25, results["Foreground IoU"]), (0.25, results["Foreground F1"]),
                 (0.25, float("nan") if fte != fte else 1.0 - min(fte / 10.0, 1.0)),
                 (0.10, results["Temporal Consistency"]), (0.10, results["Region Continuity"])]
        live = [(w, v) for (w, v) in terms if v == v]
        wsum = sum(w for w, _ in live)
        if len(live) == len(terms):
            return sum(w * v for w, v in live)
        return sum(w * v for w, v in live) / wsum if wsum > 0 else 0.0

    def get_results(self, update_best=True):
        """stream_metrics.py:140-189 (the one device->host read of the int64 counters happens here)."""
        miou, fiou, precision, recall, f1 = self._calculate_foreground_metrics(self.confusion_matrix)
        nan = float("nan")
        te, re_, fe = self.temporal_evaluator, self.region_evaluator, self.front_tracking_evaluator
        results = {
            "MIoU": miou, "Foreground IoU": fiou, "Foreground F1": f1,
            "Temporal Consistency": nan if te is None else te.get_mean_score(),
            "Front Tracking Error": nan if fe is None else fe.get_mean_error(),
            "Region Continuity": nan if re_ is None else re_.get_mean_score(),
            "Precision": precision, "Recall": recall,
        }
        if te is not None and hasattr(te, "get_detailed_statistics"):
            st = te.get_detailed_statistics()
            results.update({"Transition Accuracy": st["mean_transition"], "Stability Score": st["mean_stability"],
                            "Motion Consistency": st["mean_motion"], "Wave Segment Score": st["mean_wave_segment"]})
        if re_ is not None and hasattr(re_, "get_statistics"):
            rs = re_.get_statistics()
            if "valid_ratio" in rs:
                results["Region Valid Ratio"] = rs["valid_ratio"]
        if update_best:
            score = self._calculate_weighted_score(results)
            if score > self.best_score["weighted_score"]:
                self.best_score["weighted_score"] = score
        results["Best Score"] = self.best_score["weighted_score"]
        return results

    def to_str(self, metrics):
        """metrics/base.py:31-41."""
        string = "\n"
        for k, v in metrics.items():
            string += f"{k}: {v:.4f}\n"
        return string

    def reset(self):
        """stream_metrics.py:191-195."""
        if self._cm_dev is not None:
            self._cm_dev.zero_()
        for ev in (self.temporal_evaluator, self.region_evaluator, self.front_tracking_evaluator):
            if ev is not None:
                ev.reset()
