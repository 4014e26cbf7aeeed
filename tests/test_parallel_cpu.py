"""World-size-2 `gloo` tests (CPU) of the data-parallel host logic in iswm_b200/parallel.py.

The reference is single-process nn.DataParallel (train.py:970): per-replica BatchNorm, logits gathered
to one device, criterion over the WHOLE batch (train.py:1045-1046), gradients reduce-added. The N-rank
recipe must reproduce that: SUM-all-reduced class histogram -> global denominator,
SUM-all-reduced gradients. Here two gloo ranks run the recipe with the fp32 oracle network standing in
for the CUDA engine, and rank 0 compares against the single-process DataParallel restatement.
"""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spawn(fn, world, *args):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    out = {}
    while not q.empty():
        r, v = q.get()
        out[r] = v
    return out


def _entry(fn, rank, world, port, q, *args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, fn(rank, world, *args)))
    finally:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- bucketed all-reduce
def _bucket_job(rank, world, sizes, bucket_bytes, order):
    from iswm_b200.parallel import GradBucketer
    total = sum(sizes)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(total, generator=g)
    mine = flat.clone()
    b = GradBucketer(flat, sizes, bucket_bytes)
    launched_before_finish = 0
    for i in order:
        b.mark_ready(i)
        launched_before_finish = sum(b.launched)
    n_buckets = len(b.bounds)
    b.finish()
    # second round re-uses the bucketer (reset state)
    flat2_expected = flat.clone() * world
    for i in order:
        b.mark_ready(i)
    b.finish()
    return mine.numpy(), flat.numpy().copy(), n_buckets, launched_before_finish, b.bounds, flat2_expected.numpy()


@pytest.mark.parametrize("bucket_bytes", [64, 4096, 1 << 30])
def test_grad_bucketer_allreduce_sum_world2(bucket_bytes):
    sizes = [7, 300, 1, 64, 1000, 33, 2]
    order = list(reversed(range(len(sizes))))          # backward produces the last-registered tensors first
    out = _spawn(_bucket_job, 2, sizes, bucket_bytes, order)
    mine0, got0, nb, early, bounds, _ = out[0]
    mine1, got1, _, _, _, _ = out[1]
    # after round 1 every rank held sum; round 2 summed those again -> 2 * (a + b) on both
    np.testing.assert_allclose(got0, 2 * (mine0 + mine1), rtol=1e-6)
    np.testing.assert_allclose(got1, got0, rtol=0)
    # buckets tile the buffer exactly, last range first
    covered = sorted(bounds)
    assert covered[0][0] == 0 and covered[-1][1] == sum(sizes)
    for (s0, e0), (s1, e1) in zip(covered, covered[1:]):
        assert e0 == s1
    assert bounds[0][1] == sum(sizes)
    if bucket_bytes == 64:
        assert nb > 1 and early >= nb - 1            # buckets launch as soon as their tensors are ready
    if bucket_bytes == 1 << 30:
        assert nb == 1


def _bucket_out_of_order(rank, world):
    from iswm_b200.parallel import GradBucketer
    sizes = [10, 20, 30, 40]
    flat = torch.full((100,), float(rank + 1))
    b = GradBucketer(flat, sizes, 80)               # 20 floats per bucket
    for i in (1, 3, 0, 2):                            # arbitrary readiness order must still reduce everything once
        b.mark_ready(i)
    b.finish()
    return flat.numpy().copy()


def test_grad_bucketer_any_order():
    out = _spawn(_bucket_out_of_order, 2)
    np.testing.assert_array_equal(out[0], np.full(100, 3.0, dtype=np.float32))
    np.testing.assert_array_equal(out[1], out[0])


# ----------------------------------------------------------------------------- global-batch loss recipe
def _tiny_net(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1, bias=False), torch.nn.BatchNorm2d(8), torch.nn.ReLU(),
                               torch.nn.Conv2d(8, 2, 1))


def _batch(seed, B=4, H=12, W=10):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((B, 3, H, W), generator=g)
    y = (torch.rand((B, H, W), generator=g) < 0.3).long()
    y[0] = (torch.rand((H, W), generator=g) < 0.9).long()        # very different class mix per shard
    y[torch.rand((B, H, W), generator=g) < 0.05] = 255
    return x, y


def _dp_rank_job(rank, world):
    """The recipe of parallel.DataParallel.train_step, with torch autograd standing in for the engine."""
    from oracle import oracle_np as O
    net = _tiny_net(5).train()
    x, y = _batch(9)
    w = torch.tensor([1.0, 4.0])
    xs, ys = x.chunk(world)[rank], y.chunk(world)[rank]
    logits = net(xs)
    hist = torch.tensor(O.class_hist(ys.numpy(), 2))
    dist.all_reduce(hist, op=dist.ReduceOp.SUM)                     # parallel.py: _allreduce_hist
    D = float((hist.double() * w.double()).sum())
    _, g_local = O.weighted_ce(logits.detach().numpy(), ys.numpy(), w.numpy())
    # O.weighted_ce normalises by the LOCAL denominator; rescale to the global one
    lh = O.class_hist(ys.numpy(), 2)
    D_local = float((lh * w.numpy().astype(np.float64)).sum())
    num_local, _ = O.weighted_ce(logits.detach().numpy(), ys.numpy(), w.numpy())
    num_local = num_local * D_local
    dlogits = torch.tensor(g_local * (D_local / D), dtype=torch.float32)
    logits.backward(dlogits)
    flat = torch.cat([p.grad.flatten() for p in net.parameters()])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                     # gradient all-reduce is SUM, not mean
    num = torch.tensor([num_local], dtype=torch.float64)
    dist.all_reduce(num, op=dist.ReduceOp.SUM)
    return flat.numpy(), float(num.item() / D), hist.numpy()


def test_global_batch_loss_and_grads_match_dataparallel_semantics():
    world = 2
    out = _spawn(_dp_rank_job, world)
    # single-process restatement of nn.DataParallel (train.py:970, :1045-1048): per-replica BN forward,
    # logits gathered, ONE criterion over the whole batch, gradients summed over replicas
    x, y = _batch(9)
    w = torch.tensor([1.0, 4.0])
    replicas = [_tiny_net(5).train() for _ in range(world)]
    logits = torch.cat([net(xs) for net, xs in zip(replicas, x.chunk(world))])
    loss = torch.nn.CrossEntropyLoss(weight=w, ignore_index=255, reduction="mean")(logits, y)
    loss.backward()
    ref = sum(torch.cat([p.grad.flatten() for p in net.parameters()]) for net in replicas).numpy()
    for r in range(world):
        flat, gl, hist = out[r]
        np.testing.assert_allclose(flat, ref, rtol=2e-4, atol=1e-6)
        assert abs(gl - loss.item()) <= 1e-5 * abs(loss.item())
        assert hist.tolist() == [int((y == 0).sum()), int((y == 1).sum())]
    # and the DDP-style "mean of per-rank means" is NOT the same thing on this batch
    per_rank_mean = np.mean([torch.nn.CrossEntropyLoss(weight=w, ignore_index=255)(l, t).item()
                             for l, t in zip(logits.detach().chunk(world), y.chunk(world))])
    assert abs(per_rank_mean - loss.item()) > 1e-3 * abs(loss.item())
