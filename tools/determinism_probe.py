"""Run the same train step twice from identical state; report where results differ (logits, loss, per-parameter grads)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iswm_b200.network import modeling
from iswm_b200.utils.loss import CrossEntropyLoss
from oracle.gen_golden import seeded_state_dict, synth_labels
dev = torch.device("cuda", 0)
B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 12, 64, 64
g = torch.Generator().manual_seed(11)
x = torch.randn((B, 3, H, W), generator=g).to(dev)
y = synth_labels((B, H, W), seed=12, fg=0.3).to(dev)
w = torch.tensor([1.0, 4.0])
outs = []
for run in range(3):
    m = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False)
    m.load_state_dict(seeded_state_dict(m.state_dict(), 5))
    m.to(dev).train()
    m.engine().dropout_p = 0.0
    eng = m.engine()
    eng.debug_taps = {}
    crit = CrossEntropyLoss(weight=w).to(dev)
    logits = m(x)
    loss = crit(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    outs.append((logits.detach().clone(), float(loss), {n: p.grad.clone() for n, p in m.named_parameters()}, dict(eng.debug_taps)))
def rel(a, b): return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))
for r in (1, 2):
    print(f"run {r} vs 0: logits rel {rel(outs[r][0], outs[0][0]):.3e}  loss {outs[r][1]:.7f} vs {outs[0][1]:.7f}")
    first = [(n, rel(t, outs[0][3][n])) for n, t in outs[r][3].items()]
    bad = [(n, e) for n, e in first if e > 1e-6]
    print("  first differing forward taps:", bad[:6])
    gd = sorted(((rel(outs[r][2][n], outs[0][2][n]), n) for n in outs[0][2]), reverse=True)
    print("  worst grads:", [(n, f"{e:.2e}") for e, n in gd[:6]])
