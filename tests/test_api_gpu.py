"""The reference-facing Python objects driven on the GPU the way train.py / evaluate_quantization.py drive them:
StreamMetrics.update / get_results / reset / update_cuda / _fast_hist (metrics/stream_metrics.py:102-195),
calculate_class_weights (train.py:388-410), setup_criterion (train.py:454-459) - against the golden vectors written by
the reference's own classes (oracle/gen_golden.py) and the KATs of SURVEY.md 8c. Integer results bit-exact."""
import math
import os
import types

import numpy as np
import pytest
import torch

from iswm_b200 import metrics, train_utils
from iswm_b200.metrics import StreamMetrics, StreamSegMetrics
from iswm_b200.utils.loss import CrossEntropyLoss
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def lm(golden_dir):
    return np.load(os.path.join(golden_dir, "loss_metric.npz"))


def test_streammetrics_update_matches_reference_accumulation(lm):
    """Same call sequence as the golden generator: update(seq, sequence_data=True) uses the LAST frame only
    (stream_metrics.py:113-114), then update(frame, sequence_data=False); confusion matrix bit-exact, ratios to 1e-12."""
    assert StreamSegMetrics is StreamMetrics and metrics.StreamMetrics is StreamMetrics
    sm = StreamMetrics(2, shape_metrics=False)               # the confusion-matrix path alone; the evaluators: tests/test_shape_gpu.py
    t, p = lm["upd_true"], lm["upd_pred"]
    sm.update(t, p, sequence_data=True)
    sm.update(t[0], p[0], sequence_data=False)
    cm = sm.confusion_matrix
    assert cm.dtype == np.float64 and cm.shape == (2, 2)
    assert np.array_equal(cm, lm["upd_cm"])
    res = sm.get_results()
    got = [res["MIoU"], res["Foreground IoU"], res["Foreground F1"], res["Precision"], res["Recall"]]
    np.testing.assert_allclose(got, lm["upd_vals"], rtol=0, atol=1e-12)
    for key in ("MIoU", "Foreground IoU", "Foreground F1", "Temporal Consistency", "Front Tracking Error", "Region Continuity",
                "Precision", "Recall", "Best Score"):
        assert key in res
    assert isinstance(sm.to_str(res), str)
    # reset (stream_metrics.py:191-195): counters cleared, best score kept (the reference never clears it)
    best = sm.best_score["weighted_score"]
    sm.reset()
    assert np.array_equal(sm.confusion_matrix, np.zeros((2, 2)))
    assert sm.best_score["weighted_score"] == best > 0.0


def test_streammetrics_update_refreshes_best_score_every_call(lm):
    """stream_metrics.py:124-137: update() itself re-evaluates the running results and keeps the best weighted score and
    its components, without any get_results() call from the user."""
    sm = StreamMetrics(2, shape_metrics=False)
    t, p = lm["upd_true"], lm["upd_pred"]
    sm.update(t[0], t[0], sequence_data=False)                   # perfect prediction first
    first = dict(sm.best_score)
    assert first["weighted_score"] > 0.0 and abs(first["miou"] - 1.0) < 1e-6 and abs(first["foreground_f1"] - 1.0) < 1e-6
    sm.update(t[1], p[1], sequence_data=False)                   # a random one lowers the running score: best stays
    assert sm.best_score == first
    r = sm.get_results(update_best=False)
    assert r["Best Score"] == first["weighted_score"] and r["MIoU"] < 1.0


def test_streammetrics_plugin_evaluators_follow_the_reference_formula(lm):
    """With the three shape evaluators plugged in the weighted score is the reference's 0.05/0.25/0.25/0.25/0.10/0.10 mix
    (stream_metrics.py:65-100); without them the missing terms are NaN and the weights renormalise."""
    class Const:
        def __init__(self, v): self.v, self.n = v, 0
        def update(self, pred, gt): self.n += 1
        def reset(self): self.n = 0
        def get_mean_score(self): return self.v
        def get_mean_error(self): return self.v
    t, p = lm["upd_true"], lm["upd_pred"]
    sm = StreamMetrics(2)
    sm.temporal_evaluator, sm.region_evaluator, sm.front_tracking_evaluator = Const(0.6), Const(0.7), Const(2.5)
    sm.update(t, p, sequence_data=True)
    assert sm.temporal_evaluator.n == sm.region_evaluator.n == sm.front_tracking_evaluator.n == 1
    r = sm.get_results()
    want = 0.05 * r["MIoU"] + 0.25 * r["Foreground IoU"] + 0.25 * r["Foreground F1"] + 0.25 * (1 - 2.5 / 10) + 0.10 * 0.6 + 0.10 * 0.7
    assert abs(sm._calculate_weighted_score(r) - want) < 1e-12 and abs(r["Best Score"] - want) < 1e-12
    bare = StreamMetrics(2, shape_metrics=False)
    bare.update(t, p, sequence_data=True)
    rb = bare.get_results()
    assert math.isnan(rb["Front Tracking Error"]) and math.isnan(rb["Temporal Consistency"]) and math.isnan(rb["Region Continuity"])
    wantb = (0.05 * rb["MIoU"] + 0.25 * rb["Foreground IoU"] + 0.25 * rb["Foreground F1"]) / 0.55
    assert abs(rb["Best Score"] - wantb) < 1e-12


def test_streammetrics_fast_hist_and_update_cuda(lm):
    for name, n in (("h2", 2), ("h5", 5)):
        sm = StreamMetrics(n)
        assert np.array_equal(sm._fast_hist(lm[f"{name}_true"], lm[f"{name}_pred"]), lm[f"{name}_hist"])
        # tensor fast path: integer predictions; accumulates on the device, twice -> 2 x the histogram
        t, p = torch.tensor(lm[f"{name}_true"]).to(DEV), torch.tensor(lm[f"{name}_pred"]).to(DEV)
        sm.update_cuda(t, p)
        sm.update_cuda(t, p)
        assert np.array_equal(sm.confusion_matrix, 2.0 * lm[f"{name}_hist"])
    # logits fast path: argmax (train.py:644) and softmax[:,1] > t (evaluate_quantization.py:265-269) vs the golden class maps
    lg = torch.tensor(lm["am_logits"]).to(DEV)
    g = torch.Generator().manual_seed(3)
    y = torch.randint(0, 2, lg.shape[:1] + lg.shape[2:], generator=g)
    y[0, 0, :3] = 255
    for thr, pred in ((None, lm["am_argmax"]), (0.5, lm["am_thresh"])):
        sm = StreamMetrics(2)
        sm.update_cuda(y.to(DEV), lg, threshold=thr)
        assert np.array_equal(sm.confusion_matrix, O.fast_hist(y.numpy().reshape(-1), pred.reshape(-1), 2).astype(np.float64))


def test_calculate_class_weights_kat3_and_batch_formats(lm, capsys):
    """KAT-3: black=1000, white=37 -> [1.0, sqrt(1000/37)] as FloatTensor; tuples and {'mask':...} dicts; 255 counted in neither
    class (train.py:388-410)."""
    lab = torch.full((1, 40, 30), 255, dtype=torch.uint8)
    flat = lab.view(-1)
    flat[:1000] = 0
    flat[1000:1037] = 1
    halves = (lab[:, :20].contiguous(), lab[:, 20:].contiguous())
    loader_tuples = [(torch.zeros(1), halves[0]), (torch.zeros(1), halves[1].long())]
    loader_dicts = [{"mask": halves[0].to(torch.int32)}, {"mask": halves[1]}]
    for loader in (loader_tuples, loader_dicts):
        w = train_utils.calculate_class_weights(loader, device=DEV)
        assert isinstance(w, torch.FloatTensor) and w.dtype == torch.float32 and tuple(w.shape) == (2,)
        assert np.array_equal(w.numpy(), lm["kat3_w"])
        assert np.array_equal(w.numpy(), O.class_weights(1000, 37))
    assert "Black: 1000, White: 37" in capsys.readouterr().out
    with pytest.raises(ValueError):
        train_utils.calculate_class_weights([torch.zeros(3)], device=DEV)


def test_setup_criterion_dispatch_and_kat2(lm):
    """train.py:454-459: 'ce_loss' -> unweighted, 'IWce_loss' -> class-weighted, anything else -> None; KAT-2 through the
    object the train loop would hold."""
    w = torch.tensor([1.0, 3.0])
    assert train_utils.setup_criterion(types.SimpleNamespace(loss_type="focal_loss"), w) is None
    ce = train_utils.setup_criterion(types.SimpleNamespace(loss_type="ce_loss"), w)
    iw = train_utils.setup_criterion(types.SimpleNamespace(loss_type="IWce_loss"), w)
    assert isinstance(ce, CrossEntropyLoss) and ce.weight is None and ce.ignore_index == 255
    assert isinstance(iw, CrossEntropyLoss) and torch.equal(iw.weight.cpu(), w)
    x = torch.tensor([[[[2.0, -1.0], [0.5, 0.0]], [[0.0, 1.0], [0.5, 3.0]]]], device=DEV, requires_grad=True)
    y = torch.tensor([[[0, 1], [255, 1]]], device=DEV)
    loss = iw.to(DEV)(x, y)
    loss.backward()
    assert abs(loss.item() - float(lm["kat2_loss"])) < 1e-6
    np.testing.assert_allclose(x.grad.cpu().numpy(), lm["kat2_grad"], rtol=1e-5, atol=1e-8)


def test_cross_entropy_check_labels_flags_out_of_range_values():
    """Labels outside [0, C) other than ignore_index: the kernels treat them as ignored (torch asserts on the device); the
    opt-in check raises instead, and leaves a clean batch alone."""
    from iswm_b200.utils.loss import CrossEntropyLoss
    logits = torch.randn((2, 2, 16, 16), device=DEV)
    y = torch.randint(0, 2, (2, 16, 16), device=DEV)
    y[0, 0, 0] = 255
    crit = CrossEntropyLoss(weight=torch.tensor([1.0, 2.0]), ignore_index=255, check_labels=True).to(DEV)
    ok = crit(logits, y)
    y2 = y.clone()
    y2[1, 3, 3] = 2
    y2[1, 4, 4] = 128
    with pytest.raises(ValueError, match="2 label values outside"):
        crit(logits, y2)
    silent = CrossEntropyLoss(weight=torch.tensor([1.0, 2.0]), ignore_index=255).to(DEV)
    y3 = y.clone()
    y3[1, 3, 3] = 255
    y3[1, 4, 4] = 255
    assert torch.equal(silent(logits, y2), silent(logits, y3))       # out-of-range == ignored, as documented
    assert torch.isfinite(ok)
