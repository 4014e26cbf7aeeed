"""Pin the oracle's restatement of the SURVEY 8f rows (focal loss, input transforms, Adam/AdamW, predict arithmetic)
against vectors produced by the REAL reference classes (tests/golden/next_rows.npz, oracle/gen_golden_next.py)."""
import os

import numpy as np
import pytest

from oracle import oracle_np as O


@pytest.fixture(scope="module")
def nr(golden_dir):
    return np.load(os.path.join(golden_dir, "next_rows.npz"))


@pytest.mark.parametrize("i", range(5))
def test_focal_matches_reference(nr, i):
    a, gm, sa, w1 = nr["focal_cases"][i]
    w = None if w1 < 0 else [1.0, w1]
    loss, grad = O.focal_loss(nr["focal_logits"], nr["focal_labels"], a, gm, bool(sa), 255, w)
    assert abs(loss - float(nr[f"focal_loss_{i}"])) <= 3e-6 * max(1.0, abs(loss))
    np.testing.assert_allclose(grad, nr[f"focal_grad_{i}"], rtol=3e-4, atol=3e-7)


def test_focal_three_classes(nr):
    loss, grad = O.focal_loss(nr["focal_logits3"], nr["focal_labels3"], 0.25, 2.0, True, 255, [1.0, 2.0, 0.5])
    assert abs(loss - float(nr["focal_loss_c3"])) <= 3e-6
    np.testing.assert_allclose(grad, nr["focal_grad_c3"], rtol=3e-4, atol=3e-7)


def test_focal_gamma0_equals_scaled_ce(nr):
    """gamma == 0: focal = alpha * CE_weighted_mean * D / N (the identity the first implementation relied on)."""
    x, y = nr["focal_logits"], nr["focal_labels"]
    w = [1.0, 3.0]
    ce, _ = O.weighted_ce(x, y, w, 255)
    h = O.class_hist(y, 2)
    D = h[0] * w[0] + h[1] * w[1]
    fl, _ = O.focal_loss(x, y, 0.5, 0.0, True, 255, w)
    assert abs(fl - 0.5 * ce * D / y.size) < 1e-12


def test_to_tensor_normalize_bit_exact(nr):
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    got = O.to_tensor_normalize(nr["tf_img"], mean, std)
    assert got.dtype == np.float32 and np.array_equal(got, nr["tf_val_img"])
    assert np.array_equal(nr["tf_lbl"], nr["tf_val_lbl"])


def test_crop_flip_bit_exact(nr):
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    i, j, h, w = (int(v) for v in nr["tf_crop_ijhw"])
    assert np.array_equal(O.to_tensor_normalize(O.crop_flip(nr["tf_img"], j, i, h, w, False), mean, std), nr["tf_crop_img"])
    assert np.array_equal(O.to_tensor_normalize(O.crop_flip(nr["tf_img"], j, i, h, w, True), mean, std), nr["tf_flip_img"])
    assert np.array_equal(O.crop_flip(nr["tf_lbl"], j, i, h, w, False), nr["tf_crop_lbl"])
    assert np.array_equal(O.crop_flip(nr["tf_lbl"], j, i, h, w, True), nr["tf_flip_lbl"])


@pytest.mark.parametrize("name,dec", [("adam", False), ("adamw", True)])
@pytest.mark.parametrize("wd", [0.0, 1e-4])
def test_adam_matches_torch(nr, name, dec, wd):
    got = O.adam_steps(nr["adam_p0"], nr["adam_grads"], weight_decay=wd, decoupled=dec)
    np.testing.assert_allclose(got, nr[f"{name}_wd{wd:g}"], rtol=2e-6, atol=2e-7)


def test_predict_arithmetic(nr):
    up = O.upsample_bilinear_nchw(nr["pred_lo"], 32, 48)
    np.testing.assert_allclose(up, nr["pred_up"], rtol=1e-5, atol=1e-6)
    # threshold / confidence from the REFERENCE's upsampled logits: exact except where fp32 exp rounding differs
    for thr in (0.2, 0.5):
        pred, conf = O.threshold_pred(nr["pred_up"], thr)
        assert (pred.astype(np.uint8) != nr[f"pred_mask_{thr:g}"]).mean() <= 1e-3
    assert (np.abs(conf.astype(int) - nr["pred_conf"].astype(int)) > 1).sum() == 0
