"""One launch (after a warm-up launch) of each kernel shape that loses the most time in a cfg2 step, for an
`ncu --set full` capture: usage: python tools/prof_shapes.py [reps]   (ncu: -k regex:'conv_igemm|conv_wgrad|bn_bwd' )"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iswm_b200 import _lib, ops  # noqa: E402

L = _lib.lib()
dev = "cuda:0"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = lambda: torch.cuda.current_stream().cuda_stream


def timed(name, fn, flops=0.0, byt=0.0):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = min(ts[1:]) if len(ts) > 1 else ts[0]
    print(f"{name}: {t:.1f} us  {flops / t / 1e6:.1f} TF/s  {byt / t / 1e3:.1f} GB/s  (abort={ops.abort_code()})", flush=True)


def conv(B, Cin, H, W, Cout, k, dil, flags, tag):
    x = torch.randn((B, H, W, Cin), device=dev).to(torch.bfloat16)
    w = torch.randn((Cout, Cin, k, k), device=dev) * 0.05
    wp = ops.pack_weight_fwd(w)
    out = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
    res = torch.randn((B, H, W, Cout), device=dev).to(torch.bfloat16) if flags & _lib.EPI_RESIDUAL else None
    d = ops.make_conv_desc(B, H, W, Cin, Cin, B, H, W, Cout, Cout, ops.conv_taps(k, dil), flags=flags, res_ld=Cout)
    fl = 2.0 * B * H * W * Cout * Cin * k * k
    byt = x.numel() * 2 + out.numel() * 2 + (res.numel() * 2 if res is not None else 0)
    timed(f"conv {tag} B{B} {Cin}->{Cout} k{k} d{dil} {H}x{W} flags={flags}",
          lambda: ops.conv_igemm(d, x, wp, out, res=res, stats=stats if flags & _lib.EPI_STATS else None), fl, byt)


def wgrad(B, Cin, H, W, Cout, k, tag):
    x = torch.randn((B, H, W, Cin), device=dev).to(torch.bfloat16)
    dy = torch.randn((B, H, W, Cout), device=dev).to(torch.bfloat16)
    dw = torch.zeros((Cout, k * k, Cin), dtype=torch.float32, device=dev)
    d = ops.make_conv_desc(B, H, W, Cin, Cin, B, H, W, Cout, Cout, ops.conv_taps(k, 1))
    fl = 2.0 * B * H * W * Cout * Cin * k * k
    timed(f"wgrad {tag} B{B} {Cin}->{Cout} k{k} {H}x{W}", lambda: ops.conv_wgrad(d, x, dy, dw), fl, (x.numel() + dy.numel()) * 2)


def bn_reduce(M, C, tag):
    x = torch.randn((M, C), device=dev).to(torch.bfloat16)
    out = torch.randn((M, C), device=dev).to(torch.bfloat16)
    gm = torch.ones(C, device=dev); bt = torch.zeros(C, device=dev); save = torch.zeros(2 * C, device=dev); save[C:] = 1
    sums = torch.zeros(2 * C + 2, dtype=torch.float64, device=dev)
    timed(f"bn_bwd_reduce {tag} M={M} C={C}",
          lambda: _lib.check(L.iswm_bn_bwd_reduce(out.data_ptr(), C, x.data_ptr(), C, None, C, M, C, save.data_ptr(), save[C:].data_ptr(), gm.data_ptr(), bt.data_ptr(),
                                                  1, 0.0, 0, None, sums.data_ptr(), st()), "r"), 0.0, 4.0 * M * C)


conv(16, 64, 128, 128, 64, 3, 1, _lib.EPI_STATS, "layer1.conv2 fwd")
conv(16, 64, 128, 128, 256, 1, 1, _lib.EPI_RESIDUAL, "layer1.conv1 dgrad (+=)")
conv(16, 256, 128, 128, 64, 1, 1, _lib.EPI_STATS, "layer1.conv1 fwd")
conv(16, 64, 128, 128, 256, 1, 1, _lib.EPI_STATS, "layer1.conv3 fwd")
conv(16, 256, 32, 32, 1024, 1, 1, _lib.EPI_STATS, "layer3.conv3 fwd")
conv(16, 304, 128, 128, 256, 3, 1, _lib.EPI_STATS, "decoder.0 fwd")
wgrad(16, 64, 128, 128, 64, 3, "layer1.conv2")
wgrad(16, 304, 128, 128, 256, 3, "decoder.0")
wgrad(16, 2048, 32, 32, 256, 1, "aspp.0")
wgrad(16, 256, 128, 128, 64, 1, "layer1.conv1")
bn_reduce(16384, 1024, "layer3.conv3")
bn_reduce(262144, 256, "layer1.conv3")
