"""One train step from identical state, repeated; per-parameter gradient differences against the first repeat."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import torch
from test_graphs_gpu import _batches, _model, DEV
from iswm_b200.utils.loss import CrossEntropyLoss
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
(x, y), = _batches(1)
crit = CrossEntropyLoss(weight=torch.tensor([1.0, 3.0])).to(DEV)
ref = None
for r in range(reps):
    m = _model(); m.engine().dropout_p = float(os.environ.get("PDROP", "0.1"))
    junk = [torch.full((int(torch.randint(1, 64, (1,))) << 18,), float("nan"), device=DEV) for _ in range(4)]   # poison freed blocks
    del junk
    loss = crit(m(x), y); loss.backward()
    torch.cuda.synchronize()
    g = {n: p.grad.clone() for n, p in m.named_parameters()}
    bufs = {n: b.clone().float() for n, b in m.named_buffers()}
    if ref is None:
        ref = (g, bufs); continue
    bad = [(n, float((g[n] - ref[0][n]).norm() / (ref[0][n].norm() + 1e-30))) for n in g]
    nan = [n for n in g if not torch.isfinite(g[n]).all()]
    bad = [(n, e) for n, e in bad if not (e < 1e-5)]
    bb = [(n, float((bufs[n] - ref[1][n]).norm() / (ref[1][n].norm() + 1e-30))) for n in bufs]
    bb = [(n, e) for n, e in bb if not (e < 1e-6)]
    print(f"rep {r}: loss {float(loss):.6f}  {len(bad)} parameter grads differ, {len(nan)} non-finite, {len(bb)} buffers differ")
    if bad:
        print("   last (forward order) differing:", bad[-4:])
        print("   first differing:", bad[:3])
    if bb:
        print("   buffers:", bb[:4])
