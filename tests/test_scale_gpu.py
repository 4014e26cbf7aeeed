"""GPU parity of the device train transform (iswm_random_scale_crop through ops / DeviceTransform): bit-exact against the
fixtures the REAL reference pipeline produced (oracle/gen_golden_scale.py) and against the numpy oracle on random batches."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "scale_rows.npz"))
MEAN, STD = [float(v) for v in G["mean"]], [float(v) for v in G["std"]]


def test_scaled_pipeline_equals_the_reference_fixtures():
    from iswm_b200.data import DeviceTransform
    dev = torch.device("cuda:0")
    geom = torch.from_numpy(G["geom"].copy())
    B = geom.shape[0]
    img = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(G["img"], (B,) + G["img"].shape))).to(dev)
    lbl = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(G["lbl"], (B,) + G["lbl"].shape))).to(dev)
    tf = DeviceTransform(MEAN, STD, crop_size=tuple(int(v) for v in G["crop"]), hflip=True, scale_range=(0.5, 2.0), pad_if_needed=True)
    x, y = tf(img, lbl, params=geom)
    assert x.dtype == torch.float32 and y.dtype == torch.uint8
    assert np.array_equal(x.cpu().numpy(), G["out_img"])
    assert np.array_equal(y.cpu().numpy(), G["out_lbl"])
    # image alone (predict-style call), and a second call that reuses the table workspace
    x2 = tf(img, None, params=geom)
    assert torch.equal(x2, x)


@pytest.mark.parametrize("Hs,Ws,H,W,C,lo,hi", [(64, 80, 48, 48, 3, 0.5, 2.0), (37, 53, 40, 24, 3, 0.3, 1.2), (96, 96, 64, 64, 1, 0.5, 2.0),
                                              (512, 512, 512, 512, 3, 0.5, 2.0)])
def test_scaled_pipeline_equals_the_oracle(Hs, Ws, H, W, C, lo, hi):
    from iswm_b200 import ops
    from iswm_b200.data import DeviceTransform
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(Hs + W)
    B = 6 if Hs < 512 else 3
    img = rng.randint(0, 256, (B, Hs, Ws, C), dtype=np.uint8)
    lbl = (rng.rand(B, Hs, Ws) < 0.3).astype(np.uint8)
    lbl[rng.rand(B, Hs, Ws) < 0.02] = 255
    mean, std = MEAN[:C], STD[:C]
    tf = DeviceTransform(mean, std, crop_size=(H, W), hflip=True, scale_range=(lo, hi), pad_if_needed=True, generator=torch.Generator().manual_seed(Hs))
    scales = [lo, hi, 1.0] + list(rng.uniform(lo, hi, size=B - 3))
    geom = tf.draw_scaled(B, Hs, Ws, scales=scales)
    x, y = tf(torch.from_numpy(img).to(dev), torch.from_numpy(lbl).to(dev), params=geom)
    x, y = x.cpu().numpy(), y.cpu().numpy()
    for b in range(B):
        sh, sw, pad, y0, x0, fl = [int(v) for v in geom[b, :6]]
        assert (sh, sw, pad) == O.random_scale_geometry(Hs, Ws, float(scales[b]), (H, W))[:3]
        ri, rl = O.random_scale_crop(img[b], lbl[b], sh, sw, pad, y0, x0, H, W, bool(fl), np.float32(mean), np.float32(std))
        if C == 1:
            ri = ri.reshape(1, H, W)
        assert np.array_equal(ri, x[b]), (b, sh, sw, pad, np.abs(ri - x[b]).max())
        assert np.array_equal(rl, y[b]), (b, sh, sw, pad)
    assert ops.random_scale_kmax(Hs, Ws, geom.tolist()) >= 3


def test_scaled_pipeline_draws_are_reproducible_and_refuse_bad_input():
    from iswm_b200 import ops
    from iswm_b200.data import DeviceTransform
    dev = torch.device("cuda:0")
    img = torch.randint(0, 256, (4, 48, 64, 3), dtype=torch.uint8, device=dev)
    lbl = torch.randint(0, 2, (4, 48, 64), dtype=torch.uint8, device=dev)
    a = DeviceTransform(MEAN, STD, crop_size=32, hflip=True, scale_range=(0.5, 2.0), pad_if_needed=True, generator=torch.Generator().manual_seed(9))(img, lbl)
    b = DeviceTransform(MEAN, STD, crop_size=32, hflip=True, scale_range=(0.5, 2.0), pad_if_needed=True, generator=torch.Generator().manual_seed(9))(img, lbl)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and a[0].shape == (4, 3, 32, 32)
    e = DeviceTransform(MEAN, STD, crop_size=32, hflip=True, scale_range=(0.5, 2.0), pad_if_needed=True)(img[:0], lbl[:0])     # empty batch
    assert e[0].shape == (0, 3, 32, 32) and e[1].shape == (0, 32, 32)
    with pytest.raises(TypeError):
        ops.random_scale_crop(img.float(), lbl, torch.zeros((4, 8), dtype=torch.int32, device=dev), (32, 32), MEAN, STD, 5, (96, 128))
    with pytest.raises(TypeError):
        ops.random_scale_crop(img, lbl, torch.zeros((4, 8), dtype=torch.int32), (32, 32), MEAN, STD, 5, (96, 128))


def test_full_size_properties_of_the_scaled_pipeline():
    """Size-independent properties at cfg2's geometry (16 x 512^2): scale 1 with a full-size crop IS the unscaled pipeline
    (Pillow returns a copy; the coefficient rows degenerate to one tap of 2^22), and an exact x2 nearest upscale of the labels
    repeats every pixel twice along both axes."""
    from iswm_b200 import ops
    from iswm_b200.data import DeviceTransform
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    B, S = 16, 512
    img = torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).to(dev)
    lbl = torch.randint(0, 2, (B, S, S), generator=g, dtype=torch.uint8).to(dev)
    tf1 = DeviceTransform(MEAN, STD, crop_size=S, hflip=True, scale_range=(1.0, 1.0), pad_if_needed=True, generator=torch.Generator().manual_seed(4))
    geom = tf1.draw_scaled(B, S, S)
    assert geom[:, :5].tolist() == [[S, S, 0, 0, 0]] * B and 0 < int(geom[:, 5].sum()) < B
    x1, y1 = tf1(img, lbl, params=geom)
    x0 = ops.u8_to_f32_norm(img, MEAN, STD, (S, S), None, geom[:, 5].to(torch.uint8).to(dev))
    y0 = ops.crop_flip_u8(lbl, (S, S), None, geom[:, 5].to(torch.uint8).to(dev))
    assert torch.equal(x1, x0) and torch.equal(y1, y0)
    tf2 = DeviceTransform(MEAN, STD, crop_size=S, hflip=False, scale_range=(2.0, 2.0), pad_if_needed=True)
    geom2 = tf2.draw_scaled(B, S, S, scales=[2.0] * B)
    geom2[:, 3], geom2[:, 4] = 256, 128                       # crop origin (y0, x0) in the 1024^2 upscaled tile
    _, y2 = tf2(img, lbl, params=geom2)
    want = lbl[:, 128:384, 64:320].repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    assert torch.equal(y2, want)
