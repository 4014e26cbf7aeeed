"""Is the cfg2 train step bound by host-side launch work? Compares the time to ENQUEUE a step with the time to run it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from iswm_b200.network import modeling
from iswm_b200.optim import FusedSGD
from iswm_b200.utils.loss import CrossEntropyLoss

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False).to(dev).train()
crit = CrossEntropyLoss(weight=torch.tensor([1.0, 7.0])).to(dev)
opt = FusedSGD(model, lr=1e-3, momentum=0.9, weight_decay=1e-4)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
xd, yd = synth_batch(B, 512, 512, 0, device=dev)

def step():
    loss = crit(model(xd), yd)
    opt.zero_grad(); loss.backward(); opt.step()
    return loss
for _ in range(3):
    step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"B={B}: enqueue {1e3 * (t1 - t0):.2f} ms, until done {1e3 * (t2 - t0):.2f} ms")
N = 10
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(N): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"B={B}: {N} steps: enqueue {1e3 * (t1 - t0) / N:.2f} ms/step, total {1e3 * (t2 - t0) / N:.2f} ms/step")
