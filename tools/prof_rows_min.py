"""One launch sequence of the round-2 'next row' kernels (fused tail, shape metrics, region components) for an ncu --set full capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iswm_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, S = 16, 512
h = S // 4
lo = torch.randn((B, h, h, 2), generator=g).to(dev)
y = (torch.rand((B, S, S), generator=g) < 0.05).long().to(dev)
wts = torch.tensor([1.0, 7.0], device=dev)
dlo = torch.empty((B, h, h, 8), dtype=torch.bfloat16, device=dev)
bias = torch.zeros(2, device=dev)
scratch = torch.zeros(8200, dtype=torch.uint8, device=dev)
masks = (torch.rand((8, S, S), generator=g) < 0.02).to(torch.uint8)
masks[:, 100:400, 200:260] = 1
masks = masks.to(dev)
for _ in range(2):
    acc, hist, num = ops.tail_fwd(lo, y, wts, 255)
    ops.tail_loss(num, wts, hist, 255)
    ops.tail_bwd(acc, wts, hist, 255, None, dlo, bias, scratch)
    ops.mask_preprocess(masks)
torch.cuda.synchronize()
print("done")
