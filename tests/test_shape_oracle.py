"""CPU pinning of the shape / front / temporal oracle (oracle/shape_np.py) against fixtures produced by the REAL reference
classes (oracle/gen_golden_shape.py -> tests/golden/shape_rows.npz), and of its OpenCV restatement against cv2 itself where
cv2 is importable."""
import os

import numpy as np
import pytest

from oracle import shape_np as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "shape_rows.npz"))
KINDS = [str(k) for k in G["kinds"]]
T, H, W = [int(v) for v in G["dims"]]


def _same(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


@pytest.mark.parametrize("k", range(len(KINDS)))
def test_preprocess_fronts_presence_equal_the_reference(k):
    preds = G[f"pred_{k}"]
    for t in range(T):
        assert np.array_equal(np.asarray(S.preprocess_mask(preds[t]), np.float64), G[f"pre_{k}"][t]), (KINDS[k], t)
        fr = np.full(H, -1, np.int32)
        for i, x in S.find_front_positions(preds[t]):
            fr[i] = x
        assert np.array_equal(fr, G[f"fronts_{k}"][t]), (KINDS[k], t)
        assert bool(S.check_wave_presence(preds[t])) == bool(G[f"wave_{k}"][t])


@pytest.mark.parametrize("k", range(len(KINDS)))
def test_front_error_region_stability_motion_equal_the_reference(k):
    preds, gts = G[f"pred_{k}"], G[f"gt_{k}"]
    fte = [S.front_tracking_error(preds[t].astype(np.int64), gts[t].astype(np.int64)) for t in range(T)]
    assert _same(fte, G[f"fte_{k}"]), (fte, G[f"fte_{k}"])
    for t in range(T):
        r = S.region_metrics(preds[t], gts[t])
        row = [np.nan] * 4 if r is None else [r["fragmentation_score"], r["similarity_score"], r["num_regions"], r["final_score"]]
        assert _same(row, G[f"reg_{k}"][t]), (KINDS[k], t, row, G[f"reg_{k}"][t])
    assert _same([S.calculate_stability(preds[t], preds[t - 1]) for t in range(1, T)], G[f"stab_{k}"])
    assert _same([S.calculate_stability(preds[t], gts[t]) for t in range(1, T)], G[f"stabgt_{k}"])
    assert _same([S.calculate_motion(preds[t], preds[t - 1]) for t in range(1, T)], G[f"mot_{k}"])


@pytest.mark.parametrize("k", range(len(KINDS)))
def test_temporal_oracle_equals_the_reference_sliding_windows(k):
    preds, gts = G[f"pred_{k}"], G[f"gt_{k}"]
    L = 3
    ev = S.TemporalOracle(sequence_length=L)
    latest = []
    for i in range(T - L + 1):
        ev.update(preds[i:i + L].astype(np.int64), gts[i:i + L].astype(np.int64))
        latest.append(ev.temporal_scores[-1] if ev.temporal_scores else 0.0)
    assert _same(latest, G[f"sm_latest_{k}"]), (latest, G[f"sm_latest_{k}"])
    keys = [str(s) for s in G["result_keys"]]
    res = dict(zip(keys, G[f"sm_results_{k}"]))
    st = ev.get_detailed_statistics()
    assert res["Temporal Consistency"] == ev.get_mean_score()
    assert (res["Transition Accuracy"], res["Stability Score"], res["Motion Consistency"], res["Wave Segment Score"]) == \
        (st["mean_transition"], st["mean_stability"], st["mean_motion"], st["mean_wave_segment"])


def test_morphology_and_labelling_equal_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(5)
    k3 = cv2.getStructuringElement(cv2.MORPH_RECT, (3, 3))
    for trial in range(25):
        h, w = rng.randint(5, 70), rng.randint(5, 70)
        m = (rng.rand(h, w) < rng.choice([0.15, 0.4, 0.6])).astype(np.uint8)
        if trial % 3 == 0:
            m = cv2.dilate(m, k3, iterations=1)
        assert np.array_equal(cv2.morphologyEx(m, cv2.MORPH_CLOSE, k3), S.box_erode(S.box_dilate(m, 1), 1))
        assert np.array_equal(cv2.morphologyEx(m, cv2.MORPH_OPEN, k3), S.box_dilate(S.box_erode(m, 1), 1))
        assert np.array_equal(cv2.erode(cv2.dilate(m, np.ones((3, 3), np.uint8), iterations=3), np.ones((3, 3), np.uint8), iterations=2),
                              S.box_erode(S.box_dilate(m, 3), 2))
        n, lab, stats, _ = cv2.connectedComponentsWithStats(m)
        mine, areas, _ = S.label8(m)
        assert n - 1 == areas.size and np.array_equal(lab, mine), trial       # same components AND the same numbering
        assert np.array_equal(stats[1:, cv2.CC_STAT_AREA], areas)
