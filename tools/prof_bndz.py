"""Data gradient with / without the BatchNorm-backward epilogue (iswm_conv_igemm_bn) against the separate bn_bwd_reduce pass, per
shape: python tools/prof_bndz.py [reps]   (ISWM_B200_LIB=tools/libiswm_b200_dbg.so prints the epilogue's phase cycles)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iswm_b200 import _lib, ops  # noqa: E402

L = _lib.lib()
dev = "cuda:0"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
only = sys.argv[2] if len(sys.argv) > 2 else ""
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = lambda: torch.cuda.current_stream().cuda_stream


def timed(fn):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts[1:]) if len(ts) > 1 else ts[0]


def case(tag, B, H, W, Cv, Cu, k, dil):
    """dgrad of conv V (Cu -> Cv): GEMM K = Cv*k*k, N = Cu; the unit that owns the activation has Cu channels."""
    if only and only not in tag:
        return
    dy_ld = ((Cv + 7) // 8) * 8
    dy = torch.randn((B, H, W, dy_ld), device=dev).to(torch.bfloat16)
    w = torch.randn((Cv, Cu, k, k), device=dev) * 0.05
    wd = ops.pack_weight_dgrad(w)
    raw = torch.randn((B, H, W, Cu), device=dev).to(torch.bfloat16)
    out = torch.empty((B, H, W, Cu), dtype=torch.bfloat16, device=dev)
    gm = torch.ones(Cu, device=dev); bt = torch.zeros(Cu, device=dev); save = torch.zeros(2 * Cu, device=dev); save[Cu:] = 1
    sums = torch.zeros(2 * Cu + 2, dtype=torch.float64, device=dev)
    M = B * H * W
    taps = [(-a, -b, 0) for (a, b, _) in ops.conv_taps(k, dil)]
    d0 = ops.make_conv_desc(B, H, W, Cv, dy_ld, B, H, W, Cu, Cu, taps)
    d1 = ops.make_conv_desc(B, H, W, Cv, dy_ld, B, H, W, Cu, Cu, taps, flags=_lib.EPI_BN_DZ)
    bnd = _lib.BnDz(raw.data_ptr(), save.data_ptr(), save[Cu:].data_ptr(), gm.data_ptr(), bt.data_ptr(), sums.data_ptr())
    t_plain = timed(lambda: ops.conv_igemm(d0, dy, wd, out))
    t_fused = timed(lambda: _lib.check(L.iswm_conv_igemm_bn(C.byref(d1), dy.data_ptr(), wd.data_ptr(), out.data_ptr(), C.byref(bnd), st()), "bn"))
    t_red = timed(lambda: _lib.check(L.iswm_bn_bwd_reduce(out.data_ptr(), Cu, raw.data_ptr(), Cu, None, Cu, M, Cu, save.data_ptr(), save[Cu:].data_ptr(),
                                                          gm.data_ptr(), bt.data_ptr(), 1, 0.0, 0, None, sums.data_ptr(), st()), "r"))
    print(f"{tag:28s} plain {t_plain:7.1f} us  fused {t_fused:7.1f} us  reduce {t_red:6.1f} us  -> {t_plain + t_red - t_fused:+7.1f} us  (abort={ops.abort_code()})", flush=True)


case("layer1.conv3 dgrad", 16, 128, 128, 256, 64, 1, 1)
case("layer1.conv2 dgrad", 16, 128, 128, 64, 64, 3, 1)
case("cls dgrad", 16, 128, 128, 2, 256, 1, 1)
case("dec2 dgrad", 16, 128, 128, 256, 256, 3, 1)
case("layer2.conv3 dgrad", 16, 64, 64, 512, 128, 1, 1)
case("layer2.conv2 dgrad", 16, 64, 64, 128, 128, 3, 1)
case("layer3.conv3 dgrad", 16, 32, 32, 1024, 256, 1, 1)
case("layer3.conv2 dgrad", 16, 32, 32, 256, 256, 3, 1)
case("layer4.conv3 dgrad", 16, 32, 32, 2048, 512, 1, 1)
case("layer4.conv2 dgrad", 16, 32, 32, 512, 512, 3, 2)
