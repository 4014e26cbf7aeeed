"""Host-side staging utilities on a GPU: HostBatchPrefetcher (double-buffered side-stream H2D) and DeferredLoss."""
from __future__ import annotations

import pytest
import torch

from iswm_b200.data import HostBatchPrefetcher
from iswm_b200.train_utils import DeferredLoss

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_prefetcher_yields_every_batch_in_order_and_reuses_two_buffers():
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn((3, 3, 16, 16), generator=g).pin_memory(), torch.randint(0, 2, (3, 16, 16), generator=g).pin_memory())
               for _ in range(7)]
    pf = HostBatchPrefetcher(batches, DEV)
    ptrs = set()
    acc = []
    for i, (x, y) in enumerate(pf):
        assert x.is_cuda and x.dtype == torch.float32 and y.dtype == torch.int64
        ptrs.add(x.data_ptr())
        # consume on the current stream BEFORE asking for the next batch (double buffering contract)
        acc.append((x.sum().item(), int(y.sum().item())))
        assert torch.equal(x.cpu(), batches[i][0]) and torch.equal(y.cpu(), batches[i][1])
    assert len(acc) == 7 and len(ptrs) == 2
    assert pf.h2d_bytes == sum(b[0].numel() * 4 + b[1].numel() * 8 for b in batches)
    # dict batches (train.py's {'mask': ...} form) and uint8 labels
    d = [{"image": b[0], "mask": b[1].to(torch.uint8)} for b in batches[:3]]
    out = [(x.clone(), y.clone()) for x, y in HostBatchPrefetcher(d, DEV)]
    assert len(out) == 3 and out[2][1].dtype == torch.uint8 and torch.equal(out[2][1].cpu(), d[2]["mask"])


def test_prefetcher_overwrite_is_ordered_after_the_consumer():
    """A slow consumer kernel on batch i must see batch i's data even though batch i+2 reuses its buffer."""
    n = 1 << 22
    batches = [(torch.full((n,), float(i)).pin_memory(), torch.zeros(1, dtype=torch.int64).pin_memory()) for i in range(6)]
    sums = []
    for x, _ in HostBatchPrefetcher(batches, DEV):
        t = x
        for _ in range(20):                  # a chain of kernels reading the staged buffer
            t = t * 1.0
        sums.append(t.sum())
    torch.cuda.synchronize()
    assert [float(s) / n for s in sums] == [float(i) for i in range(6)]


def test_deferred_loss_returns_previous_step_and_flushes_last():
    dl = DeferredLoss()
    got = []
    for i in range(5):
        v = dl.push(torch.tensor(float(i) + 0.5, device=DEV))
        got.append(v)
    assert got == [None, 0.5, 1.5, 2.5, 3.5]
    assert dl.flush() == 4.5 and dl.flush() is None
