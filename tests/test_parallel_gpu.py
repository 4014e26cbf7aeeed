"""Two-rank NCCL test of iswm_b200.parallel.DataParallel on real GPUs (skipped with fewer than 2 GPUs).

Each rank trains one step on its shard; rank 0 then replays BOTH shards on its own GPU with a second model copy
(per-shard BatchNorm, loss normalised by the GLOBAL class histogram, gradients summed) — the N-rank result must
match that single-GPU restatement of nn.DataParallel's semantics (train.py:970, :1045-1048)."""
from __future__ import annotations

import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, datetime
sys.path.insert(0, os.environ["ISWM_ROOT"])
import torch, torch.distributed as dist
from iswm_b200.network import modeling
from iswm_b200.utils.loss import CrossEntropyLoss
from iswm_b200.parallel import DataParallel
from oracle.gen_golden import seeded_state_dict, synth_labels

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))

def build():
    m = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False)
    m.load_state_dict(seeded_state_dict(m.state_dict(), 5))
    m.to(dev).train()
    m.engine().dropout_p = 0.0
    return m

B, H, W = 24, 64, 64            # 12 images per shard: the ASPP pooling BatchNorm sees 12 samples, not a degenerate 2
g = torch.Generator().manual_seed(11)
x = torch.randn((B, 3, H, W), generator=g)
y = synth_labels((B, H, W), seed=12, fg=0.3)
y[: B // 2] = synth_labels((B // 2, H, W), seed=13, fg=0.85)       # very different class mix per shard
w = torch.tensor([1.0, 4.0])
xs, ys = x.chunk(world)[rank].to(dev), y.chunk(world)[rank].to(dev)

COMM = os.environ.get("ISWM_TEST_COMM", "peer")
model = build()
crit = CrossEntropyLoss(weight=w).to(dev)
dp = DataParallel(model, crit, bucket_bytes=4 << 20, comm=COMM)
assert dp.comm_mode == COMM, dp.comm_mode
loss = dp.train_step(xs, ys, optimizer=None)
torch.cuda.synchronize()
flat = model.engine().flat_g.clone()
nb = len(dp.bucketer.bounds)
# every rank holds the SAME reduced gradient bits (fixed-order sums of the peer transport; NCCL's ring order is also rank-independent)
chk = flat.double().sum().reshape(1)
lst = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(lst, chk)
assert all(float(a) == float(lst[0]) for a in lst), [float(a) for a in lst]

if rank == 0:
    # single-GPU restatement: per-shard forward/backward with the global denominator, gradients summed
    ref_grad, num = None, 0.0
    hist = torch.zeros(2, dtype=torch.int64, device=dev)
    from iswm_b200 import ops
    ops.class_hist(y.to(dev), 2, out=hist)
    for r in range(world):
        m2 = build()
        c2 = CrossEntropyLoss(weight=w).to(dev)
        c2.hist_hook = lambda h: h.copy_(hist)                       # the all-reduced (global) histogram
        l2 = c2(m2(x.chunk(world)[r].to(dev)), y.chunk(world)[r].to(dev))
        l2.backward()
        torch.cuda.synchronize()
        gr = m2.engine().flat_g.clone()
        ref_grad = gr if ref_grad is None else ref_grad + gr
        num += float(l2)
    rel = float((flat - ref_grad).norm() / ref_grad.norm())
    print(f"RESULT rel_grad={rel:.3e} loss={float(loss):.6f} ref_loss={num:.6f} buckets={nb}")
    assert rel < 1e-4, rel          # forward and activation gradients are bit-reproducible; only wgrad's fp32 split-K atomics differ
    assert abs(float(loss) - num) <= 1e-6 * abs(num), (float(loss), num)
    assert nb > 1
dist.barrier()
if COMM == "peer":
    # the data-parallel step replayed from ONE CUDA graph against the same steps launched eagerly, from the same state
    from iswm_b200.graphs import GraphedTrainStep
    from iswm_b200.optim import FusedSGD
    res = []
    for graphed in (False, True):
        m = build()
        c = CrossEntropyLoss(weight=w).to(dev)
        d = DataParallel(m, c, bucket_bytes=4 << 20, comm="peer")
        opt = FusedSGD(m, lr=1e-2, momentum=0.9, weight_decay=1e-4)
        stepper = GraphedTrainStep(m, c, opt, dp=d) if graphed else None
        losses = []
        for _ in range(2):
            l = stepper(xs, ys) if graphed else d.train_step(xs, ys, opt)
            losses.append(float(l))
        torch.cuda.synchronize()
        res.append((losses, m.engine().flat_w.clone()))
        m.engine().grad_ready_hook = None
        dist.barrier()
    (le, we), (lg, wg) = res
    relw = float((wg - we).norm() / we.norm())
    if rank == 0:
        print(f"GRAPH_DP losses eager={le} graph={lg} rel_w={relw:.3e}")
    assert abs(le[0] - lg[0]) <= 1e-6 * abs(le[0]), (le, lg)
    assert relw <= 1e-5, relw
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("comm", ["peer", "nccl"])
def test_dataparallel_two_ranks_match_per_shard_restatement(tmp_path, comm):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, ISWM_ROOT=ROOT, ISWM_TEST_COMM=comm)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "RESULT" in r.stdout
