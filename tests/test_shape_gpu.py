"""GPU parity of the shape / front / temporal evaluators: the four kernel entry points against their numpy
restatements on random and fixture masks (bit-exact integers), and the evaluator classes end to end against fixtures produced by
the REAL reference classes (oracle/gen_golden_shape.py), StreamMetrics over sliding windows included."""
import numpy as np
import pytest
import torch

from tests import shape_fakes as F
from tests.test_shape_host_cpu import G, KINDS, check_kind

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _masks(seed, N, H, W, density):
    rng = np.random.RandomState(seed)
    m = (rng.rand(N, H, W) < density).astype(np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    for n in range(N):                                        # a few solid blobs so that valid regions exist
        for _ in range(rng.randint(0, 4)):
            cy, cx, r = rng.randint(0, H), rng.randint(0, W), rng.randint(2, max(3, min(H, W) // 4))
            m[n] |= ((yy - cy) ** 2 + (xx - cx) ** 2 < r * r).astype(np.uint8)
    return m


@pytest.mark.parametrize("N,H,W,density,dtype", [(5, 64, 96, 0.05, torch.uint8), (3, 37, 53, 0.35, torch.int64), (4, 50, 41, 0.55, torch.int32),
                                                 (2, 128, 128, 0.02, torch.uint8), (1, 7, 5, 0.5, torch.uint8)])
def test_mask_preprocess_equals_the_restatement(N, H, W, density, dtype):
    from iswm_b200 import ops
    m = _masks(N * H, N, H, W, density)
    t = torch.from_numpy(m).to(dtype)
    if dtype != torch.uint8:
        t = t * 255                                           # ignore labels count as foreground (mask > 0)
    s, f, i = ops.mask_preprocess(t.to(DEV))
    rs, rf, ri = F.mask_preprocess(t)
    assert torch.equal(s.cpu(), rs) and torch.equal(f.cpu(), rf)
    assert torch.equal(i.cpu()[:, :5], ri[:, :5]), (i.cpu(), ri)


def test_mask_preprocess_on_the_fixture_frames_and_a_full_size_frame():
    from iswm_b200 import ops
    for k in range(len(KINDS)):
        t = torch.from_numpy(G[f"pred_{k}"])
        s, f, i = ops.mask_preprocess(t.to(DEV))
        rs, rf, ri = F.mask_preprocess(t)
        assert torch.equal(s.cpu(), rs) and torch.equal(f.cpu(), rf) and torch.equal(i.cpu()[:, :5], ri[:, :5]), KINDS[k]
    big = torch.from_numpy(_masks(9, 1, 512, 512, 0.01))
    s, f, i = ops.mask_preprocess(big.to(DEV))
    rs, rf, ri = F.mask_preprocess(big)
    assert torch.equal(s.cpu(), rs) and torch.equal(f.cpu(), rf) and torch.equal(i.cpu()[:, :5], ri[:, :5])
    # idempotent at full size: the chosen region survives its own close / open unless the opening cuts it
    s2, _, i2 = ops.mask_preprocess(s)
    assert int(i2[0, 3]) <= int(i[0, 3]) + 512 * 4


def test_equal_areas_go_to_the_first_label_in_opencv_order():
    """Two 6x6 squares: raster-first is the right one (row 4), block-raster-first (cv2's numbering) is the left one (row 5, same
    block row) - np.argmax over equal areas picks cv2's label 1."""
    from iswm_b200 import ops
    m = np.zeros((1, 40, 64), np.uint8)
    m[0, 4:10, 40:46] = 1
    m[0, 5:11, 10:16] = 1
    s, f, i = ops.mask_preprocess(torch.from_numpy(m).to(DEV), min_valid_area=4.0)
    assert int(i[0, 1]) == 2 and int(i[0, 2]) == 36
    assert int(s[0, 5:11, 10:16].sum()) == 36 and int(s.sum()) == 36
    rs, _, _ = F.mask_preprocess(torch.from_numpy(m), min_valid_area=4.0)
    assert torch.equal(s.cpu(), rs)


@pytest.mark.parametrize("N,H,W,density", [(4, 64, 96, 0.02), (3, 45, 38, 0.2), (2, 128, 160, 0.004)])
def test_region_components_equal_the_restatement(N, H, W, density):
    from iswm_b200 import ops
    p, g = _masks(N + H, N, H, W, density), _masks(N + W, N, H, W, density)
    p[0] = 0                                                  # an empty prediction: counts[0] == 0 is the reference's invalid case
    for dt in (torch.uint8, torch.int64):
        c, a = ops.region_components(torch.from_numpy(p).to(dt).to(DEV), torch.from_numpy(g).to(dt).to(DEV), 50, 256)
        rc, ra = F.region_components(torch.from_numpy(p), torch.from_numpy(g), 50, 256)
        assert torch.equal(c.cpu()[:, :6], rc[:, :6]), (c.cpu(), rc)
        assert torch.equal(a.cpu().sort(dim=1).values, ra.sort(dim=1).values)


def test_front_searches_equal_the_restatement():
    from iswm_b200 import ops
    rng = np.random.RandomState(4)
    N, H, W = 6, 75, 120
    fa = rng.randint(-1, W, size=(N, H)).astype(np.int32)
    fb = rng.randint(-1, W, size=(N, H)).astype(np.int32)
    fa[rng.rand(N, H) < 0.3] = -1
    fb[rng.rand(N, H) < 0.3] = -1
    fb[1] = -1                                                # B without any point
    fa[2] = -1
    fb[3] = fa[3]                                             # zero distances, ties between rows
    d2, dx = ops.front_nearest(torch.from_numpy(fa).to(DEV), torch.from_numpy(fb).to(DEV))
    rd2, rdx = F.front_nearest(torch.from_numpy(fa), torch.from_numpy(fb))
    assert torch.equal(d2.cpu(), rd2) and torch.equal(dx.cpu(), rdx)
    other = torch.from_numpy(_masks(8, N, H, W, 0.03))
    for window in (0, 1, 12, 200):
        d = ops.front_window_diff(torch.from_numpy(fa).to(DEV), other.to(DEV), window)
        assert torch.equal(d.cpu(), F.front_window_diff(torch.from_numpy(fa), other, window)), window


@pytest.mark.parametrize("k", range(len(KINDS)))
def test_evaluators_equal_the_reference_fixtures(k):
    check_kind(k)


def test_streammetrics_default_evaluators_take_device_tensors():
    from iswm_b200.metrics import StreamMetrics
    from iswm_b200.metrics.shape_metrics import FrontTrackingMetrics, RegionMetrics, TemporalMetrics
    k = KINDS.index("speckle")
    preds, gts = torch.from_numpy(G[f"pred_{k}"]).to(DEV), torch.from_numpy(G[f"gt_{k}"]).to(DEV)
    sm = StreamMetrics(2, sequence_length=3, device=DEV)
    assert isinstance(sm.temporal_evaluator, TemporalMetrics) and isinstance(sm.region_evaluator, RegionMetrics) \
        and isinstance(sm.front_tracking_evaluator, FrontTrackingMetrics)
    for i in range(preds.shape[0] - 2):
        sm.update(gts[i:i + 3], preds[i:i + 3], sequence_data=True)
    res = sm.get_results()
    keys = [str(s) for s in G["result_keys"]]
    want = dict(zip(keys, G[f"sm_results_{k}"]))
    for key in ("Temporal Consistency", "Front Tracking Error", "Region Continuity", "MIoU", "Best Score"):
        assert float(res[key]) == float(want[key]), key


def test_full_size_properties_of_the_preprocessing():
    """At 512^2: a solid rectangle away from the border survives close / open / labelling unchanged and preprocessing is idempotent
    on its own output; speckle below the 0.1 % area bar leaves an empty support."""
    from iswm_b200 import ops
    m = torch.zeros((3, 512, 512), dtype=torch.uint8)
    m[0, 100:400, 200:260] = 1
    m[1, 100:400, 200:260] = 1
    m[1, 50:60, 30:45] = 1                                    # a second region below the bar (150 px < 262 px): not "valid"
    g = torch.Generator().manual_seed(1)
    m[2] = (torch.rand((512, 512), generator=g) < 0.002).to(torch.uint8)
    s, f, i = ops.mask_preprocess(m.to(DEV))
    assert torch.equal(s[0].cpu(), m[0]) and torch.equal(s[1].cpu(), m[0])
    assert i[:, :4].cpu().tolist()[0] == [1, 1, 18000, 18000] and i[1, :4].cpu().tolist() == [2, 1, 18000, 18000]
    assert int(s[2].sum()) == 0 and int(i[2, 1]) == 0
    assert f[0, 100:400].cpu().tolist() == [200] * 300 and int(f[0, :100].max()) == -1 and int(f[0, 400:].max()) == -1
    s2, f2, i2 = ops.mask_preprocess(s)
    assert torch.equal(s2, s) and torch.equal(f2, f) and torch.equal(i2[:2, 1:4], i[:2, 1:4])
