"""cProfile of the host side of one cfg2 train step (where the ~12 ms of Python/ctypes launch work goes)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from iswm_b200.network import modeling
from iswm_b200.optim import FusedSGD
from iswm_b200.utils.loss import CrossEntropyLoss
dev = torch.device("cuda", 0)
model = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False).to(dev).train()
crit = CrossEntropyLoss(weight=torch.tensor([1.0, 7.0])).to(dev)
opt = FusedSGD(model, lr=1e-3, momentum=0.9, weight_decay=1e-4)
xd, yd = synth_batch(16, 512, 512, 0, device=dev)
def step():
    loss = crit(model(xd), yd)
    opt.zero_grad(); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5): step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
