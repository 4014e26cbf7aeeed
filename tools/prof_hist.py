"""class_hist / wce / argmax_confusion at cfg5 (16 x 1024^2, int64 labels), median of 20 launches with an L2 flush between;
run with ISWM_B200_LIB=<variant .so> to compare builds (e.g. -DISWM_HIST_UNROLL=8)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iswm_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, H, W = 16, 1024, 1024
labels = (torch.rand((B, H, W), generator=g) < 0.02).long().to(dev)
hist = ops.class_hist(labels, 2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def med(fn, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2], ts[0]
m, lo = med(lambda: ops.class_hist(labels, 2, out=hist))
print(os.environ.get("ISWM_B200_LIB", "default"), "class_hist us median %.2f min %.2f -> %.0f GB/s" % (m, lo, labels.numel() * 8 / m / 1e3))
