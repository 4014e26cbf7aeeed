"""CPU checks of the device train transform's arithmetic (ExtRandomScale + ExtRandomCrop(pad_if_needed)):
the numpy oracle against Pillow itself and against the reference-generated fixtures (oracle/gen_golden_scale.py), and the
HOST EMULATION of the CUDA kernels (the same scale_math.h functions, compiled with g++) against both - the kernels' index
arithmetic is pinned without a GPU; tests/test_scale_gpu.py then holds the real kernels to the same fixtures."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_np as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "scale_rows.npz"))
MEAN, STD = G["mean"].astype(np.float32), G["std"].astype(np.float32)


def test_oracle_resize_equals_the_reference_fixtures():
    img, lbl = G["img"], G["lbl"]
    for k, (h, w) in enumerate(G["resize_sizes"]):
        assert np.array_equal(O.pil_resize_bilinear_u8(img, int(h), int(w)), G[f"resize_img_{k}"]), (h, w)
        assert np.array_equal(O.pil_resize_nearest(lbl, int(h), int(w)), G[f"resize_lbl_{k}"]), (h, w)


def test_oracle_resize_equals_pillow_on_random_sizes():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.RandomState(0)
    for trial in range(120):
        H, W = rng.randint(3, 80), rng.randint(3, 80)
        s = rng.uniform(0.3, 2.5)
        oh, ow = max(1, int(H * s)), max(1, int(W * s))
        if trial % 11 == 0:
            ow = W
        if trial % 13 == 0:
            oh = H
        img = rng.randint(0, 256, (H, W, 3), dtype=np.uint8)
        lbl = rng.randint(0, 256, (H, W), dtype=np.uint8)
        assert np.array_equal(np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR)), O.pil_resize_bilinear_u8(img, oh, ow)), (H, W, oh, ow)
        assert np.array_equal(np.asarray(Image.fromarray(lbl).resize((ow, oh), Image.NEAREST)), O.pil_resize_nearest(lbl, oh, ow)), (H, W, oh, ow)


def test_nearest_index_is_a_running_sum_not_a_product():
    """Geometry.c adds the step output pixel by output pixel; (x + 0.5) * step lands on the other side of exact boundaries."""
    t = O.pil_nearest_table(6, 9)
    direct = ((np.arange(9) + 0.5) * (6.0 / 9)).astype(np.int64)
    assert not np.array_equal(t, direct) and t.tolist() == [0, 1, 1, 2, 2, 3, 4, 5, 5]


def test_oracle_pipeline_equals_the_reference_pipeline():
    img, lbl, (H, W) = G["img"], G["lbl"], G["crop"]
    Hs, Ws = lbl.shape
    for k, g in enumerate(G["geom"]):
        sh, sw, pad, y0, x0, fl = [int(v) for v in g[:6]]
        assert (sh, sw, pad) == O.random_scale_geometry(Hs, Ws, float(G["scales"][k]), (int(H), int(W)))[:3]
        oi, ol = O.random_scale_crop(img, lbl, sh, sw, pad, y0, x0, int(H), int(W), bool(fl), MEAN, STD)
        assert np.array_equal(oi, G["out_img"][k]), k
        assert np.array_equal(ol, G["out_lbl"][k]), k


def test_device_transform_geometry_is_the_reference_sequence():
    from iswm_b200.data import DeviceTransform
    Hs, Ws = G["lbl"].shape
    crop = tuple(int(v) for v in G["crop"])
    for k, g in enumerate(G["geom"]):
        assert DeviceTransform.scaled_geometry(Hs, Ws, float(G["scales"][k]), crop)[:3] == tuple(int(v) for v in g[:3])
    import torch
    tf = DeviceTransform(crop_size=crop, hflip=True, scale_range=(0.5, 2.0), pad_if_needed=True, generator=torch.Generator().manual_seed(5))
    geom = tf.draw_scaled(64, Hs, Ws)
    for sh, sw, pad, y0, x0, fl, _, _ in geom.tolist():
        assert Hs // 2 <= sh <= 2 * Hs and Ws // 2 <= sw <= 2 * Ws and fl in (0, 1)
        assert 0 <= y0 <= sh + 2 * pad - crop[0] and 0 <= x0 <= sw + 2 * pad - crop[1]
    with pytest.raises(ValueError):
        DeviceTransform(crop_size=crop, scale_range=(0.5, 2.0), pad_if_needed=False).draw_scaled(1, Hs, Ws, scales=[0.5])
    with pytest.raises(ValueError):
        DeviceTransform(scale_range=(0.5, 2.0))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "scale_emul.so")
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "host_emul", "scale_emul.cpp")], check=True)
    return C.CDLL(so)


def _emul_run(L, img, lbl, geom, H, W, kmax):
    B, Hs, Ws, Cc = img.shape
    tab_w, tab_h = int(geom[:, 1].max()), int(geom[:, 0].max())
    out, lo = np.zeros((B, Cc, H, W), np.float32), np.zeros((B, H, W), np.uint8)
    g = np.ascontiguousarray(geom.astype(np.int32))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.emul_random_scale_crop(p(img), p(lbl), B, Hs, Ws, Cc, p(g), kmax, tab_w, tab_h, p(MEAN), p(STD), H, W, p(out), p(lo))
    assert rc == 0, rc
    return out, lo


def test_kernel_arithmetic_host_emulation_equals_the_reference_pipeline(emul):
    from iswm_b200 import ops
    geom = G["geom"]
    B = len(geom)
    img = np.ascontiguousarray(np.broadcast_to(G["img"], (B,) + G["img"].shape))
    lbl = np.ascontiguousarray(np.broadcast_to(G["lbl"], (B,) + G["lbl"].shape))
    H, W = [int(v) for v in G["crop"]]
    kmax = ops.random_scale_kmax(lbl.shape[1], lbl.shape[2], geom.tolist())
    assert kmax == 5 and kmax == max(emul.emul_ksize_for(lbl.shape[1], int(g[0])) for g in geom)
    out, lo = _emul_run(emul, img, lbl, geom, H, W, kmax)
    assert np.array_equal(out, G["out_img"]) and np.array_equal(lo, G["out_lbl"])
    # more taps than needed change nothing (the batch maximum is what a mixed batch runs with)
    out7, lo7 = _emul_run(emul, img, lbl, geom, H, W, 9)
    assert np.array_equal(out7, out) and np.array_equal(lo7, lo)


def test_kernel_arithmetic_host_emulation_equals_the_oracle_on_random_batches(emul):
    from iswm_b200 import ops
    rng = np.random.RandomState(3)
    for trial in range(12):
        Hs, Ws, H, W, B = rng.randint(20, 70), rng.randint(20, 70), rng.randint(8, 40), rng.randint(8, 40), 3
        img = rng.randint(0, 256, (B, Hs, Ws, 3), dtype=np.uint8)
        lbl = rng.randint(0, 3, (B, Hs, Ws)).astype(np.uint8)
        geom = np.zeros((B, 8), np.int32)
        for b in range(B):
            sh, sw, pad, Hp, Wp = O.random_scale_geometry(Hs, Ws, rng.uniform(0.3, 2.2), (H, W))
            geom[b, :6] = [sh, sw, pad, rng.randint(0, Hp - H + 1), rng.randint(0, Wp - W + 1), rng.randint(0, 2)]
        out, lo = _emul_run(emul, img, lbl, geom, H, W, ops.random_scale_kmax(Hs, Ws, geom.tolist()))
        for b in range(B):
            sh, sw, pad, y0, x0, fl = [int(v) for v in geom[b, :6]]
            ri, rl = O.random_scale_crop(img[b], lbl[b], sh, sw, pad, y0, x0, H, W, bool(fl), MEAN, STD)
            assert np.array_equal(ri, out[b]) and np.array_equal(rl, lo[b]), (trial, b)


def test_kernel_arithmetic_host_emulation_equals_pillow_over_the_whole_scale_domain(emul):
    """Scales from 1/6.5 to 3 (coefficient rows of 3 .. 15 taps, the library's limit is 16): the kernels' arithmetic against Pillow
    ITSELF (resize, then numpy pad / crop / flip and the fp32 normalisation), not only against the restatement."""
    Image = pytest.importorskip("PIL.Image")
    from iswm_b200 import ops
    rng = np.random.RandomState(17)
    Hs, Ws, H, W = 53, 67, 24, 40
    img = rng.randint(0, 256, (1, Hs, Ws, 3), dtype=np.uint8)
    lbl = rng.randint(0, 4, (1, Hs, Ws)).astype(np.uint8)
    seen = set()
    for s in [1 / 6.5, 0.2, 0.26, 1 / 3.0, 0.41, 0.5, 0.66, 0.75, 0.99, 1.0, 1.01, 1.25, 1.5, 2.0, 2.37, 3.0]:
        sh, sw, pad, Hp, Wp = O.random_scale_geometry(Hs, Ws, s, (H, W))
        y0, x0, fl = rng.randint(0, Hp - H + 1), rng.randint(0, Wp - W + 1), rng.randint(0, 2)
        geom = np.array([[sh, sw, pad, y0, x0, fl, 0, 0]], np.int32)
        kmax = ops.random_scale_kmax(Hs, Ws, geom.tolist())
        seen.add(kmax)
        out, lo = _emul_run(emul, img, lbl, geom, H, W, kmax)
        ri = np.asarray(Image.fromarray(img[0]).resize((sw, sh), Image.BILINEAR))
        rl = np.asarray(Image.fromarray(lbl[0]).resize((sw, sh), Image.NEAREST))
        ri = np.pad(ri, ((pad, pad), (pad, pad), (0, 0)))[y0:y0 + H, x0:x0 + W]
        rl = np.pad(rl, ((pad, pad), (pad, pad)))[y0:y0 + H, x0:x0 + W]
        if fl:
            ri, rl = ri[:, ::-1], rl[:, ::-1]
        assert np.array_equal(O.to_tensor_normalize(ri, MEAN, STD), out[0]), (s, sh, sw, pad)
        assert np.array_equal(rl, lo[0]), (s, sh, sw, pad)
    assert min(seen) == 3 and max(seen) == 15, seen


def test_kernel_arithmetic_host_emulation_equals_the_real_reference_pipeline_on_random_draws(emul):
    """Build-container only (skipped where /root/reference is absent): 40 random draws through the reference's own
    ExtCompose([ExtRandomScale, ExtRandomCrop(pad_if_needed), ExtRandomHorizontalFlip, ExtToTensor, ExtNormalize]) (train.py:355-362)
    against the kernels' arithmetic, image and label bit for bit."""
    import sys
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    Image = pytest.importorskip("PIL.Image")
    from iswm_b200 import ops
    from oracle.gen_golden_scale import ScriptedRandom
    ref_import.install_stubs()
    m = sys.modules.get("utils")
    if m is not None and not getattr(m, "__file__", "").startswith(ref_import.REF_ROOT):
        del sys.modules["utils"]
    import utils.ext_transforms as et  # type: ignore
    sr = ScriptedRandom()
    saved = et.random
    et.random = sr
    try:
        rng = np.random.RandomState(23)
        Hs, Ws, crop = 45, 61, (28, 36)
        img = rng.randint(0, 256, (Hs, Ws, 3), dtype=np.uint8)
        lbl = rng.randint(0, 2, (Hs, Ws)).astype(np.uint8)
        pipeline = et.ExtCompose([et.ExtRandomScale((0.5, 2.0)), et.ExtRandomCrop(size=crop, pad_if_needed=True), et.ExtRandomHorizontalFlip(),
                                  et.ExtToTensor(), et.ExtNormalize(mean=[float(v) for v in MEAN], std=[float(v) for v in STD])])
        for _ in range(40):
            sr.scale, sr.frac_i, sr.frac_j, sr.coin, sr.log = float(rng.uniform(0.5, 2.0)), float(rng.rand()), float(rng.rand()), float(rng.rand()), {}
            ti, tl = pipeline(Image.fromarray(img), Image.fromarray(lbl))
            sh, sw, pad, Hp, Wp = O.random_scale_geometry(Hs, Ws, sr.scale, crop)
            geom = np.array([[sh, sw, pad, sr.log.get("i", 0), sr.log.get("j", 0), int(sr.coin < 0.5), 0, 0]], np.int32)
            out, lo = _emul_run(emul, img[None], lbl[None], geom, crop[0], crop[1], ops.random_scale_kmax(Hs, Ws, geom.tolist()))
            assert np.array_equal(out[0], ti.numpy()), (sr.scale, geom)
            assert np.array_equal(lo[0], tl.numpy().astype(np.uint8)), (sr.scale, geom)
    finally:
        et.random = saved
