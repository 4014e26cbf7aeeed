"""GPU parity of the SURVEY 8f rows (callers either side of the hot path) against the oracle and the reference golden
vectors: fused focal loss, device input pipeline, fused Adam/AdamW, fused predict epilogue."""
import os

import numpy as np
import pytest
import torch

from iswm_b200 import ops
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


@pytest.fixture(scope="module")
def nr(golden_dir):
    return np.load(os.path.join(golden_dir, "next_rows.npz"))


# ----------------------------------------------------------------------------- focal loss
@pytest.mark.parametrize("i", range(5))
def test_focal_vs_reference_golden(nr, i):
    from iswm_b200.utils.loss import FocalLoss
    a, gm, sa, w1 = nr["focal_cases"][i]
    w = None if w1 < 0 else torch.tensor([1.0, float(w1)])
    x = torch.from_numpy(nr["focal_logits"]).to(DEV).requires_grad_(True)
    y = torch.from_numpy(nr["focal_labels"]).to(DEV)
    crit = FocalLoss(alpha=float(a), gamma=float(gm), size_average=bool(sa), ignore_index=255, weight=w).to(DEV)
    loss = crit(x, y)
    loss.backward()
    ref = float(nr[f"focal_loss_{i}"])
    assert abs(float(loss.detach()) - ref) <= 1e-5 * max(1.0, abs(ref))          # fp32 kernel vs fp32 torch
    np.testing.assert_allclose(x.grad.cpu().numpy(), nr[f"focal_grad_{i}"], rtol=2e-3, atol=2e-6)


def test_focal_three_classes_and_label_dtypes(nr):
    for dt in (torch.int64, torch.int32, torch.uint8):
        loss, grad = ops.focal_fwd_bwd(torch.from_numpy(nr["focal_logits3"]).to(DEV), torch.from_numpy(nr["focal_labels3"]).to(dt).to(DEV),
                                       torch.tensor([1.0, 2.0, 0.5]), 0.25, 2.0, True, 255)
        assert abs(float(loss) - float(nr["focal_loss_c3"])) <= 1e-6
        np.testing.assert_allclose(grad.cpu().numpy(), nr["focal_grad_c3"], rtol=2e-3, atol=2e-6)


@pytest.mark.parametrize("gamma,sa", [(0.0, True), (2.0, True), (1.0, False)])
def test_focal_vs_oracle_ragged(gamma, sa):
    g = torch.Generator().manual_seed(5)
    x = torch.randn((3, 2, 37, 53), generator=g) * 3
    y = (torch.rand((3, 37, 53), generator=g) < 0.1).long()
    y[torch.rand((3, 37, 53), generator=g) < 0.05] = 255
    w = torch.tensor([1.0, 7.0])
    loss, grad = ops.focal_fwd_bwd(x.to(DEV), y.to(DEV), w, 0.75, gamma, sa, 255)
    rl, rg = O.focal_loss(x.numpy(), y.numpy(), 0.75, gamma, sa, 255, w.numpy())
    assert abs(float(loss) - rl) <= 2e-5 * max(1.0, abs(rl))
    np.testing.assert_allclose(grad.cpu().numpy(), rg, rtol=2e-3, atol=1e-6 * (1 if sa else x[0, 0].numel()))


def test_focal_all_ignored_and_bf16():
    x = torch.randn((1, 2, 8, 8))
    y = torch.full((1, 8, 8), 255)
    loss, grad = ops.focal_fwd_bwd(x.to(DEV), y.to(DEV), None, 1.0, 2.0, True, 255)
    assert float(loss) == 0.0 and float(grad.abs().max()) == 0.0        # mean over all pixels of zeros (not nan, unlike CE)
    y = (torch.rand((1, 8, 8)) < 0.5).long()
    xb = x.to(torch.bfloat16)
    loss, grad = ops.focal_fwd_bwd(xb.to(DEV), y.to(DEV), None, 1.0, 2.0, True, 255)
    rl, rg = O.focal_loss(xb.float().numpy(), y.numpy(), 1.0, 2.0, True, 255)
    assert grad.dtype == torch.bfloat16 and abs(float(loss) - rl) < 1e-5
    np.testing.assert_allclose(grad.float().cpu().numpy(), rg, rtol=1e-2, atol=1e-5)


# ----------------------------------------------------------------------------- device input pipeline
def test_to_tensor_normalize_bit_exact_vs_reference(nr):
    img = torch.from_numpy(nr["tf_img"])[None].to(DEV)
    out = ops.u8_to_f32_norm(img, MEAN, STD)
    assert np.array_equal(out[0].cpu().numpy(), nr["tf_val_img"])


def test_crop_flip_bit_exact_vs_reference(nr):
    i, j, h, w = (int(v) for v in nr["tf_crop_ijhw"])
    img = torch.from_numpy(np.stack([nr["tf_img"], nr["tf_img"]])).to(DEV)
    lbl = torch.from_numpy(np.stack([nr["tf_lbl"], nr["tf_lbl"]])).to(DEV)
    org = torch.tensor([[j, i], [j, i]], dtype=torch.int32, device=DEV)
    flip = torch.tensor([0, 1], dtype=torch.uint8, device=DEV)
    x = ops.u8_to_f32_norm(img, MEAN, STD, (h, w), org, flip).cpu().numpy()
    y = ops.crop_flip_u8(lbl, (h, w), org, flip).cpu().numpy()
    assert np.array_equal(x[0], nr["tf_crop_img"]) and np.array_equal(x[1], nr["tf_flip_img"])
    assert np.array_equal(y[0], nr["tf_crop_lbl"]) and np.array_equal(y[1], nr["tf_flip_lbl"])


@pytest.mark.parametrize("shape", [(1, 5, 7, 3), (3, 33, 130, 3), (2, 16, 16, 1), (2, 9, 21, 4)])
def test_normalize_ragged_vs_oracle(shape):
    rng = np.random.RandomState(3)
    img = rng.randint(0, 256, size=shape, dtype=np.uint8)
    C = shape[-1]
    mean, std = MEAN[:C] + [0.5] * (C - 3), STD[:C] + [0.25] * (C - 3)
    out = ops.u8_to_f32_norm(torch.from_numpy(img).to(DEV), mean, std).cpu().numpy()
    for b in range(shape[0]):
        assert np.array_equal(out[b], O.to_tensor_normalize(img[b], mean, std))


def test_device_transform_train_pipeline():
    from iswm_b200.data import DeviceTransform
    rng = np.random.RandomState(11)
    img = rng.randint(0, 256, size=(4, 40, 56, 3), dtype=np.uint8)
    lbl = (rng.rand(4, 40, 56) < 0.1).astype(np.uint8)
    tf = DeviceTransform(MEAN, STD, crop_size=32, hflip=True, generator=torch.Generator().manual_seed(0))
    org, flip = tf.draw(4, 40, 56)
    x, y = tf(torch.from_numpy(img).to(DEV), torch.from_numpy(lbl).to(DEV), params=(org, flip))
    assert x.shape == (4, 3, 32, 32) and y.shape == (4, 32, 32) and y.dtype == torch.uint8
    for b in range(4):
        x0, y0, f = int(org[b, 0]), int(org[b, 1]), bool(flip[b])
        assert np.array_equal(x[b].cpu().numpy(), O.to_tensor_normalize(O.crop_flip(img[b], x0, y0, 32, 32, f), MEAN, STD))
        assert np.array_equal(y[b].cpu().numpy(), O.crop_flip(lbl[b], x0, y0, 32, 32, f))
    with pytest.raises(RuntimeError):
        ops.u8_to_f32_norm(torch.from_numpy(img).to(DEV), MEAN, STD, (64, 64))      # window larger than the tile


# ----------------------------------------------------------------------------- Adam / AdamW
@pytest.mark.parametrize("name,adamw", [("adam", 0), ("adamw", 1)])
@pytest.mark.parametrize("wd", [0.0, 1e-4])
def test_adam_step_vs_torch_golden(nr, name, adamw, wd):
    from iswm_b200 import _lib
    p = torch.from_numpy(nr["adam_p0"]).to(DEV).clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for t, g in enumerate(nr["adam_grads"], 1):
        gd = torch.from_numpy(g).to(DEV)
        _lib.check(_lib.lib().iswm_adam_step(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 1e-3, 0.9, 0.999, 1e-8, wd, adamw, t, None, None,
                                             torch.cuda.current_stream().cuda_stream), "adam_step")
    np.testing.assert_allclose(p.cpu().numpy(), nr[f"{name}_wd{wd:g}"], rtol=2e-6, atol=2e-7)


def test_adam_ragged_length_vs_torch():
    from iswm_b200 import _lib
    for n in (1, 3, 4, 1027):
        g0 = torch.Generator().manual_seed(n)
        p0, gr = torch.randn(n, generator=g0), torch.randn(n, generator=g0)
        ref = torch.nn.Parameter(p0.clone())
        opt = torch.optim.AdamW([ref], weight_decay=1e-2)
        ref.grad = gr.clone()
        opt.step()
        buf = torch.zeros(((n + 3) // 4) * 4 * 4, device=DEV)             # 16-byte aligned slabs
        p, g, m, v = (buf[i * ((n + 3) // 4) * 4: i * ((n + 3) // 4) * 4 + n] for i in range(4))
        p.copy_(p0); g.copy_(gr)
        _lib.check(_lib.lib().iswm_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 1e-2, 1, 1, None, None,
                                             torch.cuda.current_stream().cuda_stream), "adam_step")
        np.testing.assert_allclose(p.cpu().numpy(), ref.detach().numpy(), rtol=2e-6, atol=2e-7)


@pytest.mark.parametrize("which", ["adam", "adamw"])
def test_fused_adam_on_model_vs_torch(which):
    """Two optimiser steps on the real parameter set: FusedAdam(W) on the engine's flat buffers vs torch.optim on copies."""
    from iswm_b200.network import modeling
    from iswm_b200.optim import FusedAdam, FusedAdamW
    torch.manual_seed(0)
    model = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False).to(DEV).train()
    eng = model.engine()
    eng.device = torch.device(DEV)
    eng._ensure_grad_buffers()
    params = eng._param_list()
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in params]
    opt = (FusedAdam if which == "adam" else FusedAdamW)(model, weight_decay=1e-4)
    ropt = (torch.optim.Adam if which == "adam" else torch.optim.AdamW)(ref_params, weight_decay=1e-4)
    for step in range(2):
        eng.flat_g.normal_(generator=None)
        for p, r in zip(params, ref_params):
            p.grad = eng.grad_views[id(p)]
            r.grad = p.grad.detach().clone()
        opt.step()
        ropt.step()
    worst = max(float((p.detach() - r.detach()).abs().max()) for p, r in zip(params, ref_params))
    assert worst <= 5e-6, worst


# ----------------------------------------------------------------------------- fused predict epilogue
def _lowres(B, h, w, seed, scale=1.5):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn((B, h, w, 2), generator=g) * scale).to(DEV)


@pytest.mark.parametrize("B,h,w,up", [(2, 8, 12, 4), (1, 5, 7, 4), (3, 16, 33, 4), (1, 32, 32, 2)])
@pytest.mark.parametrize("mode,thr", [(1, 0.5), (1, 0.2), (0, 0.5)])
@pytest.mark.parametrize("ldt", [torch.int64, torch.uint8])
def test_predict_epilogue_equals_unfused_path(B, h, w, up, mode, thr, ldt):
    """bit-identical to iswm_logits_up_fwd + iswm_argmax_confusion (maps and counts)."""
    from iswm_b200 import _lib
    lo = _lowres(B, h, w, B * 100 + h)
    H, W = h * up, w * up
    g = torch.Generator().manual_seed(9)
    y = (torch.rand((B, H, W), generator=g) < 0.3).long()
    y[torch.rand((B, H, W), generator=g) < 0.05] = 255
    y = y.to(ldt).to(DEV)
    logits = torch.empty((B, 2, H, W), device=DEV)
    _lib.check(_lib.lib().iswm_logits_up_fwd(lo.data_ptr(), B, h, w, 2, H, W, logits.data_ptr(), torch.cuda.current_stream().cuda_stream), "up")
    cm0, p0, c0 = ops.argmax_confusion(logits, y, mode=mode, threshold=thr, want_pred=True, want_conf=True)
    cm1, p1, c1 = ops.predict_epilogue(lo, H, W, y, mode=mode, threshold=thr)
    assert torch.equal(p0, p1) and torch.equal(c0, c1) and cm0.tolist() == cm1.tolist()
    _, p2, c2 = ops.predict_epilogue(lo, H, W, None, mode=mode, threshold=thr)
    assert torch.equal(p2, p1) and torch.equal(c2, c1)
    # and against the oracle chain (numpy bilinear -> softmax threshold): exact up to fp32 contraction order at the threshold
    upn = O.upsample_bilinear_nchw(lo.permute(0, 3, 1, 2).cpu().numpy(), H, W)
    pred = O.argmax_pred(upn) if mode == 0 else O.threshold_pred(upn, thr)[0]
    assert (pred.astype(np.uint8) != p1.cpu().numpy()).mean() <= 2e-3


def test_predict_epilogue_vs_reference_golden(nr):
    lo = torch.from_numpy(nr["pred_lo"]).permute(0, 2, 3, 1).contiguous().to(DEV)
    for thr in (0.2, 0.5):
        _, pred, conf = ops.predict_epilogue(lo, 32, 48, None, mode=1, threshold=thr)
        assert (pred.cpu().numpy() != nr[f"pred_mask_{thr:g}"]).mean() <= 2e-3
    assert (np.abs(conf.cpu().numpy().astype(int) - nr["pred_conf"].astype(int)) > 1).sum() == 0


def test_predict_mask_fused_equals_unfused_on_model():
    from iswm_b200.metrics import StreamMetrics
    from iswm_b200.network import modeling
    from iswm_b200.predict import predict_mask
    torch.manual_seed(0)
    model = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False).to(DEV).eval()
    x = torch.randn(2, 3, 128, 128, device=DEV)
    y = (torch.rand(2, 128, 128, device=DEV) < 0.3).to(torch.uint8)
    m0, m1 = StreamMetrics(2), StreamMetrics(2)
    p0, c0 = predict_mask(model, x, 0.5, y, m0, fused=False)
    p1, c1 = predict_mask(model, x, 0.5, y, m1, fused=True)
    assert torch.equal(p0, p1) and torch.equal(c0, c1)
    assert np.array_equal(m0.confusion_matrix, m1.confusion_matrix) and m0.confusion_matrix.sum() == 2 * 128 * 128


def test_predict_epilogue_rejects_unsupported():
    with pytest.raises(RuntimeError):
        ops.predict_epilogue(torch.zeros((1, 4, 4, 3), device=DEV), 16, 16)        # C != 2
    with pytest.raises(RuntimeError):
        ops.predict_epilogue(torch.zeros((1, 4, 5, 2), device=DEV), 16, 18)        # Wo % 4 != 0
