"""Host-side mirrors of the train.py helpers that define the hot path's semantics."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .utils.loss import CrossEntropyLoss


def calculate_class_weights(loader, device="cuda"):
    """train.py:388-410. One pass over the loader counting label==0 / label==1 pixels (exact integers,
    GPU histogram kernel, a single device->host read at the end), then `[1.0, sqrt(black/white)]`.
    Accepts `(images, labels)` tuples or `{'mask': labels}` dicts like the reference."""
    hist = torch.zeros(2, dtype=torch.int64, device=device)
    for batch in loader:
        if isinstance(batch, dict):
            labels = batch["mask"]
        elif isinstance(batch, (list, tuple)) and len(batch) == 2:
            _, labels = batch
        else:
            raise ValueError(f"Unexpected batch format: {type(batch)}")
        if labels.dtype not in (torch.uint8, torch.int32, torch.int64):
            labels = labels.long()
        ops.class_hist(labels.to(device, non_blocking=True), 2, out=hist)
    black_pixels, white_pixels = (int(v) for v in hist.cpu().tolist())
    weight_black = 1.0
    weight_white = np.sqrt(black_pixels / white_pixels)
    print(f"Pixel distribution - Black: {black_pixels}, White: {white_pixels}")
    print(f"Class weights - Black: {weight_black}, White: {weight_white}")
    return torch.FloatTensor([weight_black, weight_white])


def setup_criterion(opts, class_weights):
    """train.py:454-459 (returns None for any other loss_type, like the reference)."""
    if opts.loss_type == "ce_loss":
        return CrossEntropyLoss(ignore_index=255, reduction="mean")
    elif opts.loss_type == "IWce_loss":
        return CrossEntropyLoss(weight=class_weights, ignore_index=255, reduction="mean")
    return None


class DeferredLoss:
    """Per-step loss logging without stalling the launch pipeline.

    The reference reads the loss synchronously right after every step (`np_loss = loss.detach().cpu().numpy()`,
    train.py:1051), which idles the GPU while the host enqueues the next step. `push(loss)` starts an asynchronous
    device->host copy of THIS step's loss into pinned memory and returns the PREVIOUS step's value as a Python float
    (None on the first call): every step's loss still reaches the host, one step later; `flush()` returns the last one.
    """

    def __init__(self):
        self._host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self._ev = [None, None]
        self._n = 0

    def push(self, loss: torch.Tensor):
        slot = self._n & 1
        prev = self.flush() if self._n > 0 else None
        self._host[slot].copy_(loss.detach(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._ev[slot] = ev
        self._n += 1
        self._pending = slot
        return prev

    def flush(self):
        slot = getattr(self, "_pending", None)
        if slot is None:
            return None
        self._ev[slot].synchronize()
        self._pending = None
        return float(self._host[slot])
