"""bn_bwd_reduce over the tensor sizes of a cfg2 step: python tools/prof_bnred.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iswm_b200 import _lib
L = _lib.lib()
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = lambda: torch.cuda.current_stream().cuda_stream
for (M, C) in [(262144, 64), (65536, 128), (16384, 256), (16384, 512), (262144, 256), (65536, 512), (16384, 1024), (16384, 2048), (262144, 48)]:
    x = torch.randn((M, C), device=dev).to(torch.bfloat16)
    dout = torch.randn((M, C), device=dev).to(torch.bfloat16)
    gm = torch.ones(C, device=dev); bt = torch.zeros(C, device=dev); save = torch.zeros(2 * C, device=dev); save[C:] = 1
    sums = torch.zeros(2 * C + 2, dtype=torch.float64, device=dev)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.iswm_bn_bwd_reduce(dout.data_ptr(), C, x.data_ptr(), C, None, C, M, C, save.data_ptr(), save[C:].data_ptr(), gm.data_ptr(), bt.data_ptr(),
                                        1, 0.0, 0, None, sums.data_ptr(), st()), "r")
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    t = min(ts[1:])
    print(f"bn_bwd_reduce M={M} C={C}: {t:.1f} us  {4.0 * M * C / t / 1e3:.0f} GB/s")
