"""Run one convolution shape through the tcgen05 implicit-GEMM kernel a few times (for ncu / timing).
python tools/prof_conv.py B Cin H W Cout k dil [flags] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iswm_b200 import _lib, ops  # noqa: E402

B, Cin, H, W, Cout, k, dil = (int(v) for v in sys.argv[1:8])
flags = int(sys.argv[8]) if len(sys.argv) > 8 else _lib.EPI_STATS
reps = int(sys.argv[9]) if len(sys.argv) > 9 else 5
dev = "cuda:0"
x = torch.randn((B, H, W, Cin), device=dev).to(torch.bfloat16)
w = torch.randn((Cout, Cin, k, k), device=dev) * 0.05
wp = ops.pack_weight_fwd(w)
out = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=dev)
stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
res = torch.randn((B, H, W, Cout), device=dev).to(torch.bfloat16) if flags & _lib.EPI_RESIDUAL else None
d = ops.make_conv_desc(B, H, W, Cin, Cin, B, H, W, Cout, Cout, ops.conv_taps(k, dil), flags=flags, res_ld=Cout)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
times = []
for i in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv_igemm(d, x, wp, out, res=res, stats=stats if flags & _lib.EPI_STATS else None)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) * 1e3)
fl = 2.0 * B * H * W * Cout * Cin * k * k
byt = x.numel() * 2 + out.numel() * 2 + (res.numel() * 2 if res is not None else 0)
t = min(times[1:])
print(f"conv B{B} {Cin}->{Cout} k{k} d{dil} {H}x{W} flags={flags}: {t:.1f} us  {fl / t / 1e6:.1f} TF/s  {byt / t / 1e3:.1f} GB/s  (abort={ops.abort_code()})")
