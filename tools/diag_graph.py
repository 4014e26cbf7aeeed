"""Eager vs CUDA-graph train steps, state compared after every step (debugging aid for iswm_b200.graphs)."""
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
from test_graphs_gpu import _batches, _model, DEV
from iswm_b200.graphs import GraphedTrainStep
from iswm_b200.optim import FusedSGD
from iswm_b200.utils.loss import CrossEntropyLoss
p_drop = float(os.environ.get("PDROP", "0.1"))
same = os.environ.get("SAME", "0") == "1"
bs = _batches(3)
if same:
    bs = [bs[0]] * 3
crit = CrossEntropyLoss(weight=torch.tensor([1.0, 3.0])).to(DEV)
res = {}
for mode in ("eager", "graph"):
    m = _model(); m.engine().dropout_p = p_drop
    opt = FusedSGD(m, lr=1e-2, momentum=0.9, weight_decay=1e-4)
    st = GraphedTrainStep(m, crit, opt) if mode == "graph" else None
    eng = m.engine(); log = []
    for x, y in bs:
        if st is not None:
            l = float(st(x, y))
        else:
            loss = crit(m(x), y); opt.zero_grad(); loss.backward(); opt.step(); l = float(loss.detach())
        log.append((l, eng.flat_w.clone(), eng.flat_g.clone(), opt._mom.clone(), int(eng._step_dev.item()),
                    [b.detach().clone().float() for b in m.buffers()]))
    res[mode] = log
for i, (e, g) in enumerate(zip(res["eager"], res["graph"])):
    rd = lambda a, b: float((a - b).norm() / (a.norm() + 1e-30))
    bdiff = max(rd(a, b) for a, b in zip(e[5], g[5]))
    print(f"step {i + 1}: loss {e[0]:.6f} {g[0]:.6f} | w {rd(e[1], g[1]):.2e} grad {rd(e[2], g[2]):.2e} mom {rd(e[3], g[3]):.2e} | step_dev {e[4]} {g[4]} | buffers {bdiff:.2e}")
