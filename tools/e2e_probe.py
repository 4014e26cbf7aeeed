"""Times the cfg2 train step fed from pinned host memory: plain synchronous-order copies vs HostBatchPrefetcher."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from iswm_b200.network import modeling
from iswm_b200.optim import FusedSGD
from iswm_b200.utils.loss import CrossEntropyLoss
from iswm_b200.data import HostBatchPrefetcher

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False).to(dev).train()
crit = CrossEntropyLoss(weight=torch.tensor([1.0, 7.0])).to(dev)
opt = FusedSGD(model, lr=1e-3, momentum=0.9, weight_decay=1e-4)
xh, yh = synth_batch(16, 512, 512, 0, pinned=True)
xd, yd = xh.to(dev), yh.to(dev)
print("pinned:", xh.is_pinned(), yh.is_pinned())

def step(x, y):
    loss = crit(model(x), y)
    opt.zero_grad(); loss.backward(); opt.step()
    return loss

for _ in range(3):
    step(xd, yd)
torch.cuda.synchronize()
N = 10
def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / N * 1e3:.2f} ms/step")
timed("device-resident, no sync", lambda: [step(xd, yd) for _ in range(N)])
timed("device-resident, loss read each step", lambda: [float(step(xd, yd).detach()) for _ in range(N)])
timed("plain .to + loss read", lambda: [float(step(xh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True)).detach()) for _ in range(N)])
def pf():
    for xs, ys in HostBatchPrefetcher(((xh, yh) for _ in range(N)), dev):
        float(step(xs, ys).detach())
timed("prefetcher + loss read", pf)
timed("prefetcher + loss read (2nd)", pf)
def copy_only():
    for _ in range(N):
        xh.to(dev, non_blocking=True); yh.to(dev, non_blocking=True)
timed("H2D copies only", copy_only)
