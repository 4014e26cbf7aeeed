"""Fused train tail against the unfused kernel chain at cfg2's geometry, kernel by kernel (CUDA events, L2 flushed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
out = bench.run_next_rows(dev, bench.peaks())
for k, v in out.items():
    print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items() if a != "workload"})
from iswm_b200 import ops
g = torch.Generator().manual_seed(0)
B, S = 16, 512
h = S // 4
lo = torch.randn((B, h, h, 2), generator=g).to(dev)
y = (torch.rand((B, S, S), generator=g) < 0.05).long().to(dev)
wts = torch.tensor([1.0, 7.0], device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=10):
    fn(); ts = []
    for _ in range(n):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
acc, hist, num = ops.tail_fwd(lo, y, wts, 255)
dlo = torch.empty((B, h, h, 8), dtype=torch.bfloat16, device=dev); bias = torch.zeros(2, device=dev); scratch = torch.zeros(8200, dtype=torch.uint8, device=dev)
print("tail_fwd us", t(lambda: ops.tail_fwd(lo, y, wts, 255)), "(incl. 2 torch.zeros fills + empty_like)")
print("tail_loss us", t(lambda: ops.tail_loss(num, wts, hist, 255)))
print("tail_bwd us", t(lambda: ops.tail_bwd(acc, wts, hist, 255, None, dlo, bias, scratch)))
y8 = y.to(torch.uint8)
print("tail_fwd uint8 labels us", t(lambda: ops.tail_fwd(lo, y8, wts, 255)))
