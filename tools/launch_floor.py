"""Back-to-back launch cost of small kernels (PDL on/off via ISWM_PDL): conv (1 tile .. 1 wave), BN apply."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iswm_b200 import _lib, ops
L = _lib.lib()
dev = "cuda:0"
st = lambda: torch.cuda.current_stream().cuda_stream

def bench(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

print("ISWM_PDL =", os.environ.get("ISWM_PDL", "1"))
for (B, Cin, H, W, Cout, k, flags) in [(1, 64, 8, 16, 64, 1, 0), (1, 64, 8, 16, 64, 1, 8), (16, 256, 32, 32, 64, 1, 8), (16, 256, 32, 32, 1024, 1, 8), (16, 1024, 32, 32, 256, 1, 8), (16, 256, 32, 32, 256, 3, 8)]:
    x = torch.randn((B, H, W, Cin), device=dev).to(torch.bfloat16)
    w = torch.randn((Cout, Cin, k, k), device=dev) * 0.05
    wp = ops.pack_weight_fwd(w)
    out = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
    d = ops.make_conv_desc(B, H, W, Cin, Cin, B, H, W, Cout, Cout, ops.conv_taps(k, 1), flags=flags)
    t = bench(lambda: ops.conv_igemm(d, x, wp, out, stats=stats if flags & 8 else None))
    fl = 2.0 * B * H * W * Cout * Cin * k * k
    print(f"conv B{B} {Cin}->{Cout} k{k} {H}x{W} flags={flags}: {t:6.2f} us/launch back-to-back ({fl / t / 1e6:7.1f} TF/s)")
for (M, C) in [(128, 64), (16384, 256), (16384, 1024)]:
    x = torch.randn((M, C), device=dev).to(torch.bfloat16); out = torch.empty_like(x)
    stats = torch.stack([x.double().sum(0), (x.double() ** 2).sum(0)]).reshape(-1).contiguous()
    g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev); save = torch.empty(2 * C, device=dev)
    t = bench(lambda: _lib.check(L.iswm_bn_train_apply(x.data_ptr(), C, stats.data_ptr(), 1, M, C, g.data_ptr(), b.data_ptr(), 1e-5, 0.1, None, None, None,
                                                       save.data_ptr(), save[C:].data_ptr(), None, C, 1, 0.0, 0, None, out.data_ptr(), C, None, st())))
    print(f"bn_train_apply M={M} C={C}: {t:6.2f} us/launch back-to-back ({M * C * 4 / t / 1e3:6.0f} GB/s)")
