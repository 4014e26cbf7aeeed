"""GPU parity of the HBM-bound trio against the numpy oracle and the reference golden vectors."""
import os

import numpy as np
import pytest
import torch

from iswm_b200 import ops
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _labels(shape, seed, n_classes=2, fg=0.02, ign=0.01, dtype=torch.int64):
    g = torch.Generator().manual_seed(seed)
    if n_classes == 2:
        y = (torch.rand(shape, generator=g) < fg).long()
    else:
        y = torch.randint(0, n_classes, shape, generator=g)
    y[torch.rand(shape, generator=g) < ign] = 255
    return y.to(dtype)


@pytest.mark.parametrize("dtype", [torch.int64, torch.uint8, torch.int32])
@pytest.mark.parametrize("n", [0, 1, 7, 1000, 16 * 64 * 64 + 3])
def test_class_hist_bit_exact(dtype, n):
    y = _labels((n,), 1, dtype=dtype)
    h = ops.class_hist(y.to(DEV), 2)
    assert h.cpu().tolist() == O.class_hist(y.numpy(), 2).tolist()
    b, w = O.class_pixel_counts(y.numpy())
    assert h.cpu().tolist() == [b, w]


def test_class_hist_many_classes_and_accumulate():
    y = _labels((50001,), 2, n_classes=21)
    out = torch.zeros(21, dtype=torch.int64, device=DEV)
    ops.class_hist(y.to(DEV), 21, out=out)
    ops.class_hist(y.to(DEV), 21, out=out)
    assert out.cpu().tolist() == (2 * O.class_hist(y.numpy(), 21)).tolist()


def test_class_hist_unaligned_view():
    y = _labels((4099,), 3)
    yd = y.to(DEV)[3:]
    assert ops.class_hist(yd, 2).cpu().tolist() == O.class_hist(y.numpy()[3:], 2).tolist()


@pytest.mark.parametrize("case", ["small_w", "small_nw", "c3_w", "c3_nw", "mid_w", "mid_nw"])
def test_wce_golden(golden_dir, case):
    g = np.load(os.path.join(golden_dir, "loss_metric.npz"))
    x = torch.tensor(g[f"ce_{case}_logits"]).to(DEV)
    y = torch.tensor(g[f"ce_{case}_labels"]).to(DEV)
    w = torch.tensor(g[f"ce_{case}_weight"]).to(DEV) if case.endswith("_w") else None
    hist = ops.class_hist(y, x.shape[1])
    loss, grad = ops.wce_fwd_bwd(x, y, w, hist)
    ref = float(g[f"ce_{case}_loss"])
    assert abs(loss.item() - ref) <= 1e-5 * max(1.0, abs(ref))          # north_star: <= 1e-3 relative
    np.testing.assert_allclose(grad.cpu().numpy(), g[f"ce_{case}_grad"], rtol=1e-4, atol=1e-7)


def test_wce_kat2():
    x = torch.tensor([[[[2.0, -1.0], [0.5, 0.0]], [[0.0, 1.0], [0.5, 3.0]]]], device=DEV)
    y = torch.tensor([[[0, 1], [255, 1]]], device=DEV)
    w = torch.tensor([1.0, 3.0], device=DEV)
    hist = ops.class_hist(y, 2)
    assert hist.cpu().tolist() == [1, 2]                                  # denominator 1 + 3 + 3 = 7
    loss, grad = ops.wce_fwd_bwd(x, y, w, hist)
    assert abs(loss.item() - 0.09335345) < 1e-6
    ref = [-0.01702900, 0.05108697, 0, 0.02032538, 0.01702899, -0.05108699, 0, -0.02032537]
    np.testing.assert_allclose(grad.cpu().numpy().ravel(), ref, atol=2e-7)


@pytest.mark.parametrize("ldtype", [torch.int64, torch.uint8])
@pytest.mark.parametrize("shape", [(2, 2, 64, 64), (3, 2, 33, 17), (2, 5, 16, 24)])
def test_wce_vs_oracle(shape, ldtype):
    B, C, H, W = shape
    g = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=g) * 4
    y = _labels((B, H, W), 6, n_classes=C, fg=0.05, ign=0.03, dtype=ldtype)
    w = torch.rand(C, generator=g) * 6 + 0.5
    hist = ops.class_hist(y.to(DEV), C)
    loss, grad = ops.wce_fwd_bwd(x.to(DEV), y.to(DEV), w.to(DEV), hist)
    rl, rg = O.weighted_ce(x.numpy(), y.numpy(), w.numpy())
    assert abs(loss.item() - rl) <= 1e-5 * abs(rl)
    np.testing.assert_allclose(grad.cpu().numpy(), rg, rtol=2e-4, atol=1e-8)


def test_wce_bf16_logits():
    g = torch.Generator().manual_seed(9)
    x = (torch.randn((2, 2, 32, 64), generator=g) * 3).to(torch.bfloat16)
    y = _labels((2, 32, 64), 10, fg=0.1)
    hist = ops.class_hist(y.to(DEV), 2)
    loss, grad = ops.wce_fwd_bwd(x.to(DEV), y.to(DEV), None, hist)
    rl, rg = O.weighted_ce(x.float().numpy(), y.numpy(), None)
    assert abs(loss.item() - rl) <= 1e-5 * abs(rl)
    np.testing.assert_allclose(grad.float().cpu().numpy(), rg, rtol=1e-2, atol=1e-8)   # bf16 output rounding


def test_wce_all_ignored_is_nan_with_zero_grad():
    x = torch.randn((1, 2, 8, 8), device=DEV)
    y = torch.full((1, 8, 8), 255, device=DEV)
    hist = ops.class_hist(y, 2)
    loss, grad = ops.wce_fwd_bwd(x, y, None, hist)
    assert torch.isnan(loss).item() and torch.count_nonzero(grad).item() == 0


def test_wce_empty():
    x = torch.zeros((0, 2, 4, 4), device=DEV)
    y = torch.zeros((0, 4, 4), dtype=torch.long, device=DEV)
    loss, grad = ops.wce_fwd_bwd(x, y, None, ops.class_hist(y, 2))
    assert torch.isnan(loss).item() and grad.numel() == 0


@pytest.mark.parametrize("tdtype,pdtype", [(torch.int64, torch.int64), (torch.uint8, torch.uint8),
                                            (torch.int64, torch.uint8), (torch.int32, torch.int64)])
@pytest.mark.parametrize("n", [0, 5, 4096, 100003])
def test_confusion_bit_exact(tdtype, pdtype, n):
    t = _labels((n,), 11, fg=0.3, ign=0.05, dtype=tdtype)
    p = _labels((n,), 12, fg=0.3, ign=0.0, dtype=pdtype)
    cm = ops.confusion(t.to(DEV), p.to(DEV), 2).cpu().numpy()
    assert cm[:4].reshape(2, 2).tolist() == O.fast_hist(t.numpy(), p.numpy(), 2).tolist()
    assert cm[4] == 0


def test_confusion_kat1_and_golden(golden_dir):
    gt = torch.tensor([[0, 0, 1, 1], [1, 0, 255, 1]], device=DEV)
    pr = torch.tensor([[0, 1, 1, 0], [1, 0, 1, 1]], device=DEV)
    assert ops.confusion(gt, pr, 2).cpu().numpy()[:4].reshape(2, 2).tolist() == [[2, 1], [1, 3]]
    g = np.load(os.path.join(golden_dir, "loss_metric.npz"))
    for name, n in (("h2", 2), ("h5", 5)):
        cm = ops.confusion(torch.tensor(g[f"{name}_true"]).to(DEV), torch.tensor(g[f"{name}_pred"]).to(DEV), n)
        assert np.array_equal(cm.cpu().numpy()[: n * n].reshape(n, n), g[f"{name}_hist"])


def test_confusion_accumulates_and_is_linear():
    t = _labels((30000,), 13, fg=0.4, ign=0.02)
    p = _labels((30000,), 14, fg=0.4, ign=0.0)
    whole = ops.confusion(t.to(DEV), p.to(DEV), 2)
    parts = torch.zeros(5, dtype=torch.int64, device=DEV)
    ops.confusion(t[:12345].to(DEV), p[:12345].to(DEV), 2, out=parts)
    ops.confusion(t[12345:].to(DEV), p[12345:].to(DEV), 2, out=parts)
    assert torch.equal(whole, parts)
    assert int(whole[:4].sum()) == int(((t >= 0) & (t < 2)).sum())


def test_argmax_confusion_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss_metric.npz"))
    lg = torch.tensor(g["am_logits"]).to(DEV)
    t = torch.tensor(g["am_argmax"]).to(DEV)            # use the reference argmax as "truth": cm must be diagonal
    cm, pred, _ = ops.argmax_confusion(lg, t, mode=0, want_pred=True)
    assert np.array_equal(pred.cpu().numpy(), g["am_argmax"])
    c = cm.cpu().numpy()[:4].reshape(2, 2)
    assert c[0, 1] == 0 and c[1, 0] == 0
    cm2, pred2, conf2 = ops.argmax_confusion(lg, t, mode=1, threshold=0.5, want_pred=True, want_conf=True)
    assert np.array_equal(pred2.cpu().numpy(), g["am_thresh"])
    assert np.max(np.abs(conf2.cpu().numpy().astype(int) - g["am_conf"].astype(int))) <= 1
    assert pred.cpu().numpy()[0, 0, 0] == 0 and pred2.cpu().numpy()[0, 0, 0] == 0      # KAT-4 tie -> class 0


@pytest.mark.parametrize("shape", [(2, 2, 64, 64), (2, 2, 31, 9), (2, 4, 16, 16)])
@pytest.mark.parametrize("ldtype", [torch.float32, torch.bfloat16])
def test_argmax_confusion_vs_oracle(shape, ldtype):
    B, C, H, W = shape
    g = torch.Generator().manual_seed(21)
    x = torch.randn(shape, generator=g).to(ldtype)
    t = _labels((B, H, W), 22, n_classes=C, fg=0.3, ign=0.05)
    cm, pred, _ = ops.argmax_confusion(x.to(DEV), t.to(DEV), mode=0, want_pred=True)
    rp = O.argmax_pred(x.float().numpy())
    assert np.array_equal(pred.cpu().numpy(), rp)
    assert np.array_equal(cm.cpu().numpy()[: C * C].reshape(C, C), O.fast_hist(t.numpy(), rp, C))


def test_full_size_properties():
    """BASELINE cfg5 size (16x2x1024x1024): size-independent properties instead of a CPU oracle pass."""
    B, H, W = 16, 1024, 1024
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn((B, 2, H, W), device=DEV, generator=g)
    u = torch.rand((B, H, W), device=DEV, generator=g)
    y = (u < 0.02).long()
    y[u > 0.99] = 255
    hist = ops.class_hist(y, 2)
    assert hist.cpu().tolist() == [int((y == 0).sum()), int((y == 1).sum())]
    w = torch.tensor([1.0, 7.0], device=DEV)
    loss, grad = ops.wce_fwd_bwd(x, y, w, hist)
    # gradient of a softmax CE sums to zero over classes at every pixel, and is zero where ignored
    assert float((grad[:, 0] + grad[:, 1]).abs().max()) < 1e-12
    assert torch.count_nonzero(grad[:, 0][y == 255]).item() == 0
    ref = torch.nn.functional.cross_entropy(x, y, weight=w, ignore_index=255)
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    cm, pred, _ = ops.argmax_confusion(x, y, mode=0, want_pred=True)
    assert int(cm[:4].sum()) == int(hist.sum())                      # every non-ignored pixel lands in one cell
    assert cm[:4].view(2, 2).sum(1).cpu().tolist() == hist.cpu().tolist()   # row sums = class histogram
    cm2 = ops.confusion(y, pred, 2)
    assert torch.equal(cm, cm2)
