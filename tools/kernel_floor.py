"""True device cost per launch of the BatchNorm / small conv kernels: N back-to-back launches captured in a CUDA graph
(no host work between them), replayed and timed. usage: python tools/kernel_floor.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iswm_b200 import _lib, ops
L = _lib.lib()
dev = torch.device("cuda:0")
N = 40


def graph_time(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * N)


st = lambda: torch.cuda.current_stream().cuda_stream
print("ISWM_PDL =", os.environ.get("ISWM_PDL", "1"))
for (M, C) in [(16, 256), (16384, 256), (16384, 1024), (65536, 512), (262144, 256), (1048576, 64)]:
    x = torch.randn((M, C), device=dev).to(torch.bfloat16)
    out = torch.empty_like(x)
    dy = torch.empty_like(x)
    stats = torch.stack([x.double().sum(0), (x.double() ** 2).sum(0)]).reshape(-1).contiguous()
    gm = torch.ones(C, device=dev); bt = torch.zeros(C, device=dev); save = torch.empty(2 * C, device=dev)
    sums = torch.zeros(2 * C + 2, dtype=torch.float64, device=dev)
    dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    fa = lambda: _lib.check(L.iswm_bn_train_apply(x.data_ptr(), C, stats.data_ptr(), 1, M, C, gm.data_ptr(), bt.data_ptr(), 1e-5, 0.1, None, None, None,
                                                  save.data_ptr(), save[C:].data_ptr(), None, C, 1, 0.0, 0, None, out.data_ptr(), C, None, st()), "a")
    fr = lambda: _lib.check(L.iswm_bn_bwd_reduce(out.data_ptr(), C, x.data_ptr(), C, None, C, M, C, save.data_ptr(), save[C:].data_ptr(), gm.data_ptr(), bt.data_ptr(),
                                                 1, 0.0, 0, None, sums.data_ptr(), st()), "r")
    fb = lambda: _lib.check(L.iswm_bn_bwd_apply(out.data_ptr(), C, x.data_ptr(), C, None, C, M, C, gm.data_ptr(), bt.data_ptr(), save.data_ptr(), save[C:].data_ptr(),
                                                sums.data_ptr(), 1, 0.0, 0, None, dy.data_ptr(), C, None, 0, dg.data_ptr(), db.data_ptr(), st()), "b")
    ta, tr, tb = graph_time(fa), graph_time(fr), graph_time(fb)
    mb = M * C * 2 / 1e6
    print(f"M={M:8d} C={C:5d} ({mb:6.1f} MB/tensor): apply {ta:6.2f} us ({2 * mb / ta * 1e3 / 1e3:6.0f} GB/s)  bwd_reduce {tr:6.2f} us ({2 * mb / tr:6.0f} GB/s)  bwd_apply {tb:6.2f} us ({3 * mb / tb:6.0f} GB/s)")

for (B, Cin, H, W, Cout, k, flags) in [(1, 64, 8, 16, 64, 1, 0), (1, 64, 8, 16, 64, 1, 8), (16, 256, 32, 32, 64, 1, 8), (16, 256, 32, 32, 1024, 1, 8), (16, 1024, 32, 32, 256, 1, 8),
                                        (16, 256, 32, 32, 256, 3, 8), (16, 64, 128, 128, 256, 1, 8), (16, 256, 128, 128, 64, 1, 8), (16, 64, 128, 128, 64, 3, 8)]:
    x = torch.randn((B, H, W, Cin), device=dev).to(torch.bfloat16)
    w = torch.randn((Cout, Cin, k, k), device=dev) * 0.05
    wp = ops.pack_weight_fwd(w)
    out = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
    d = ops.make_conv_desc(B, H, W, Cin, Cin, B, H, W, Cout, Cout, ops.conv_taps(k, 1), flags=flags)
    t = graph_time(lambda: ops.conv_igemm(d, x, wp, out, stats=stats if flags & 8 else None))
    fl = 2.0 * B * H * W * Cout * Cin * k * k
    by = 2.0 * B * H * W * (Cin + Cout)
    print(f"conv B{B} {Cin}->{Cout} k{k} {H}x{W} flags={flags}: {t:6.2f} us/launch ({fl / t / 1e6:7.1f} TF/s, {by / t / 1e3:6.0f} GB/s)")
