"""Times the three BatchNorm kernels at the (M, C) shapes of R50-OS16 B=16 512^2 with an L2 flush between launches.
python tools/bn_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iswm_b200 import _lib
L = _lib.lib()
dev = "cuda:0"
st = lambda: torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [(1048576, 64), (262144, 64), (262144, 256), (65536, 128), (65536, 512), (16384, 256), (16384, 512), (16384, 1024), (16384, 2048), (262144, 48)]

def timeit(fn, reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]

tot = [0.0, 0.0, 0.0]
for M, C in shapes:
    x = torch.randn((M, C), device=dev).to(torch.bfloat16)
    res = torch.randn((M, C), device=dev).to(torch.bfloat16)
    out = torch.empty_like(x); dy = torch.empty_like(x); dz = torch.empty_like(x)
    dout = torch.randn((M, C), device=dev).to(torch.bfloat16)
    stats = torch.stack([x.double().sum(0), (x.double() ** 2).sum(0)]).reshape(-1).contiguous()
    g = torch.rand(C, device=dev) + 0.5; b = torch.randn(C, device=dev) * 0.1
    rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev); nbt = torch.zeros((), dtype=torch.long, device=dev)
    save = torch.empty(2 * C, device=dev); sums = torch.zeros(2 * C, dtype=torch.float64, device=dev); dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    for with_res in (False, True):
        rp = res.data_ptr() if with_res else None
        t1 = timeit(lambda: _lib.check(L.iswm_bn_train_apply(x.data_ptr(), C, stats.data_ptr(), 1, M, C, g.data_ptr(), b.data_ptr(), 1e-5, 0.1, rm.data_ptr(), rv.data_ptr(), nbt.data_ptr(),
                                                             save.data_ptr(), save[C:].data_ptr(), rp, C, 1, 0.0, 0, None, out.data_ptr(), C, None, st())))
        ap = out.data_ptr() if with_res else None
        t2 = timeit(lambda: _lib.check(L.iswm_bn_bwd_reduce(dout.data_ptr(), C, x.data_ptr(), C, ap, C, M, C, save.data_ptr(), save[C:].data_ptr(), g.data_ptr(), b.data_ptr(), 1, 0.0, 0, None, sums.data_ptr(), st())))
        t3 = timeit(lambda: _lib.check(L.iswm_bn_bwd_apply(dout.data_ptr(), C, x.data_ptr(), C, ap, C, M, C, g.data_ptr(), b.data_ptr(), save.data_ptr(), save[C:].data_ptr(), sums.data_ptr(), 1, 0.0, 0, None,
                                                           dy.data_ptr(), C, dz.data_ptr() if with_res else None, C, dg.data_ptr(), db.data_ptr(), st())))
        e = M * C * 2
        b1, b2, b3 = e * (3 if with_res else 2), e * (3 if with_res else 2), e * (5 if with_res else 3)
        print(f"M={M:8d} C={C:5d} res={int(with_res)}  apply {t1:7.1f} us {b1 / t1 / 1e3:6.0f} GB/s | reduce {t2:7.1f} us {b2 / t2 / 1e3:6.0f} GB/s | bwd_apply {t3:7.1f} us {b3 / t3 / 1e3:6.0f} GB/s")
