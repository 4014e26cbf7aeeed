"""Is the eager train step paced by the host? Times forward + loss + backward (+ SGD) of cfg2 eagerly and as a CUDA-graph
replay (same kernels, zero host work per launch). usage: python tools/graph_probe.py [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iswm_b200.network import modeling
from iswm_b200.optim import FusedSGD
from iswm_b200.utils.loss import CrossEntropyLoss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False).to(dev).train()
crit = CrossEntropyLoss(weight=torch.tensor([1.0, 7.0])).to(dev)
opt = FusedSGD(model, lr=1e-3, momentum=0.9, weight_decay=1e-4)
x = torch.randn(B, 3, 512, 512, device=dev)
y = (torch.rand(B, 512, 512, device=dev) < 0.02).long()


def step():
    loss = crit(model(x), y)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


def timed(fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host


for _ in range(3):
    step()
ms, host = timed(step)
print(f"eager: {ms:.2f} ms/step on the device, {host:.2f} ms/step of host enqueue time")
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    loss = step()
torch.cuda.synchronize()
for _ in range(3):
    g.replay()
ms, host = timed(g.replay)
print(f"graph: {ms:.2f} ms/step on the device, {host:.3f} ms/step of host time; loss {float(loss.detach()):.5f}")

# marginal cost of each kernel family inside the pipeline: the same step captured with that family's launches skipped
from iswm_b200 import _lib
eng = model.engine()
for inline in (False, True):
    eng.async_wgrad = not inline
    res = {}
    for name, mask in (("all", 0), ("no conv_igemm", 1), ("no conv_wgrad", 2), ("no BatchNorm kernels", 4), ("no conv at all", 3), ("only glue + loss + SGD", 7)):
        _lib.lib().iswm_debug_set_skip(mask)
        gg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gg):
            step()
        torch.cuda.synchronize()
        for _ in range(2):
            gg.replay()
        res[name], _ = timed(gg.replay, 5)
        _lib.lib().iswm_debug_set_skip(0)
    print(f"weight gradients {'in line' if inline else 'on the side stream'}: " + "; ".join(f"{k} {v:.2f} ms" for k, v in res.items()))

# forward branches on the side stream (Engine._fwd_fork) on / off, same process, graph replays
eng.async_wgrad = True
for ov in (False, True, False, True):
    eng.fwd_overlap = ov
    gg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gg):
        step()
    torch.cuda.synchronize()
    for _ in range(2):
        gg.replay()
    t, _ = timed(gg.replay, 8)
    print(f"forward branch overlap {'on ' if ov else 'off'}: {t:.3f} ms/step")
