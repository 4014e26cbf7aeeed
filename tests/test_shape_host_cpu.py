"""CPU checks of the host half of the shape / front / temporal evaluators (iswm_b200/metrics/shape_metrics.py): with the four
kernel entry points replaced by numpy stand-ins (tests/shape_fakes.py) every score must equal the fixtures the REAL reference
classes produced (oracle/gen_golden_shape.py) - bit for bit, including StreamMetrics driven over sliding windows as
train.py:676-681 does."""
import os

import numpy as np
import pytest

from tests import shape_fakes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "shape_rows.npz"))
KINDS = [str(k) for k in G["kinds"]]
T, H, W = [int(v) for v in G["dims"]]


def same(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


def check_kind(k):
    """Shared by the CPU (fakes) and GPU (kernels) tests."""
    from iswm_b200.metrics import StreamMetrics
    from iswm_b200.metrics.shape_metrics import FrontTrackingMetrics, MaskUtils, RegionMetrics
    preds, gts = G[f"pred_{k}"], G[f"gt_{k}"]
    for t in range(T):
        assert np.array_equal(np.asarray(MaskUtils.preprocess_mask(preds[t]), np.float64), G[f"pre_{k}"][t]), (KINDS[k], t)
        fr = np.full(H, -1, np.int32)
        for i, x in MaskUtils.find_front_positions(preds[t]):
            fr[i] = x
        assert np.array_equal(fr, G[f"fronts_{k}"][t]), (KINDS[k], t)
        assert MaskUtils.check_wave_presence(preds[t], 0.005) == bool(G[f"wave_{k}"][t])
    fte = [FrontTrackingMetrics().update(preds[t].astype(np.int64), gts[t].astype(np.int64)) for t in range(T)]
    assert same(fte, G[f"fte_{k}"]), (KINDS[k], fte, G[f"fte_{k}"])
    for t in range(T):
        r = RegionMetrics().calculate_region_metrics(preds[t].astype(np.int64), gts[t].astype(np.int64))
        row = [np.nan] * 4 if r is None else [r["fragmentation_score"], r["similarity_score"], r["num_regions"], r["final_score"]]
        assert same(row, G[f"reg_{k}"][t]), (KINDS[k], t, row, G[f"reg_{k}"][t])
    assert same([MaskUtils.calculate_stability(preds[t], preds[t - 1]) for t in range(1, T)], G[f"stab_{k}"])
    assert same([MaskUtils.calculate_stability(preds[t], gts[t]) for t in range(1, T)], G[f"stabgt_{k}"])
    assert same([MaskUtils.calculate_motion(preds[t], preds[t - 1]) for t in range(1, T)], G[f"mot_{k}"])
    # the validation loop of train.py:676-681
    L = 3
    sm = StreamMetrics(2, sequence_length=L, device=MaskUtils.device)
    latest = []
    for i in range(T - L + 1):
        sm.update(gts[i:i + L].astype(np.int64), preds[i:i + L].astype(np.int64), sequence_data=True)
        latest.append(sm.temporal_evaluator.get_latest_score())
    assert same(latest, G[f"sm_latest_{k}"]), (KINDS[k], latest, G[f"sm_latest_{k}"])
    res = sm.get_results()
    keys = [str(s) for s in G["result_keys"]]
    got = [float(res.get(key, np.nan)) for key in keys]
    assert same(got, G[f"sm_results_{k}"]), (KINDS[k], dict(zip(keys, zip(got, G[f"sm_results_{k}"]))))
    assert np.array_equal(sm.confusion_matrix, G[f"sm_cm_{k}"]) and sm.best_score["weighted_score"] == float(G[f"sm_best_{k}"][0])
    sm.reset()
    assert sm.temporal_evaluator.temporal_scores == [] and sm.region_evaluator.total_cases == 0 and sm.front_tracking_evaluator.tracking_errors == []


@pytest.fixture
def fakes(monkeypatch):
    import torch
    from iswm_b200 import ops
    shape_fakes.install(monkeypatch)

    def confusion(t, p, n, out=None):                       # numpy stand-in for the confusion kernel (StreamMetrics.update)
        from oracle import oracle_np as O
        cm = torch.from_numpy(np.concatenate([O.fast_hist(t.numpy(), p.numpy(), n).reshape(-1), [0]]).astype(np.int64))
        if out is None:
            return cm
        out += cm
        return out
    monkeypatch.setattr(ops, "confusion", confusion)


@pytest.mark.parametrize("k", range(len(KINDS)))
def test_host_half_equals_the_reference_fixtures(fakes, k):
    check_kind(k)


def test_weighted_masks_hold_no_pixel_equal_to_one(fakes):
    """mask_utils.py:38-43 returns base_mask * weight for several valid regions: the `== 1` scans of find_front_positions /
    calculate_stability then see nothing - restated, not 'fixed'."""
    from iswm_b200.metrics.shape_metrics import MaskUtils
    k = KINDS.index("twins")
    pre = MaskUtils.preprocess(G[f"pred_{k}"][0])
    assert pre.valid == 2 and pre.weight == 0.8 and pre.count > 0 and pre.front_positions() == []
    assert set(np.unique(pre.numpy()).tolist()) == {0.0, 0.8}


def test_host_half_equals_the_real_reference_classes_at_the_default_window(fakes):
    """Build-container only (skipped where /root/reference or cv2 is absent): StreamMetrics with the DEFAULT sequence_length=7 over
    20 frames of every kind, driven exactly as train.py:676-681 does - the product's evaluators (on the numpy stand-ins for the
    kernels) against the reference's own classes running here, every reported number bit for bit."""
    import contextlib
    import io
    pytest.importorskip("cv2")
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    from oracle.gen_golden_shape import make_frames
    _, ref_metrics = ref_import.reference_modules()
    from iswm_b200.metrics import StreamMetrics
    L, Tn = 7, 20
    for k, kind in enumerate(KINDS):
        preds, gts = make_frames(300 + k, Tn, 48, 80, kind)
        ref = ref_metrics.StreamMetrics(2, sequence_length=L)
        mine = StreamMetrics(2, sequence_length=L, device="cpu")
        with contextlib.redirect_stdout(io.StringIO()):
            for i in range(Tn - L + 1):
                ref.update(gts[i:i + L].astype(np.int64), preds[i:i + L].astype(np.int64), sequence_data=True)
                mine.update(gts[i:i + L].astype(np.int64), preds[i:i + L].astype(np.int64), sequence_data=True)
                assert same([ref.temporal_evaluator.get_latest_score()], [mine.temporal_evaluator.get_latest_score()]), (kind, i)
            a, b = ref.get_results(), mine.get_results()
        assert set(a) == set(b), (kind, set(a) ^ set(b))
        for key in a:
            assert same([float(a[key])], [float(b[key])]), (kind, key, a[key], b[key])
        assert np.array_equal(ref.confusion_matrix, mine.confusion_matrix)
        assert same(ref.front_tracking_evaluator.tracking_errors, mine.front_tracking_evaluator.tracking_errors), kind
        assert same(ref.region_evaluator.valid_scores, mine.region_evaluator.valid_scores), kind


def test_wave_presence_of_a_weighted_mask_at_the_threshold(fakes):
    """A weighted mask whose ratio sits ON the threshold: the decision is numpy's own sum of the float mask (mask_utils.py:139-142)."""
    from iswm_b200.metrics.shape_metrics import MaskUtils
    from oracle import shape_np as S
    m = np.zeros((100, 100), np.uint8)
    m[10:40, 10:31] = 1                                      # 630 px, the largest region
    m[60:80, 60:80] = 1                                      # 400 px: a second valid region -> weight 0.8
    pre = MaskUtils.preprocess(m)
    assert pre.valid == 2 and pre.count == 630
    ref_mask = S.preprocess_mask(m)
    exact = float(np.sum(ref_mask) / ref_mask.size)
    for thr in (exact, np.nextafter(exact, 1.0), np.nextafter(exact, 0.0), 0.8 * 630 / 10000, 0.05, 0.051):
        assert MaskUtils.check_wave_presence(m, thr) == bool(S.check_wave_presence(m, thr)), thr
