"""Data parallelism for the hot path: one process per GPU, NCCL over NVLink 5 / NVSwitch.

The reference is single-process nn.DataParallel (train.py:970): parameters replicated, the batch
scattered along dim 0, PER-REPLICA BatchNorm statistics, logits gathered to GPU 0 where the criterion
normalises over the WHOLE batch (train.py:1045-1046), gradients reduce-added to GPU 0. The same
semantics here, without the GPU-0 bottleneck:
  * each rank runs forward/backward on its shard with its own BN statistics;
  * the per-class pixel histogram is all-reduced (SUM, C int64) before the fused CE kernel, so every
    rank divides by the GLOBAL denominator sum_c w_c n_c; gradients are then all-reduced with SUM
    (a DDP-style mean of per-rank means would differ whenever class mixes differ across ranks);
  * gradient all-reduce is bucketed over the engine's flat fp32 gradient buffer and launched from the
    backward sweep as soon as the last tensor of a bucket has been produced (reverse registration
    order: head first, stem last), overlapping NCCL with the remaining backward kernels.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class GradBucketer:
    """Splits a flat gradient buffer into ~bucket_bytes contiguous buckets aligned to tensor
    boundaries and all-reduces each one when all of its tensors have been marked ready.
    Pure torch.distributed logic (works with gloo on CPU for tests, NCCL on GPUs)."""

    def __init__(self, flat: torch.Tensor, sizes: List[int], bucket_bytes: int = 25 << 20, group=None):
        self.flat, self.group = flat, group
        self.bounds = []                 # (start, end) element ranges, in registration order
        self.tensor_bucket = []
        elems = max(1, bucket_bytes // flat.element_size())
        # build from the END (backward produces the last-registered parameters first)
        cuts, acc, off = [], 0, sum(sizes)
        ends = []
        o = 0
        for n in sizes:
            o += n
            ends.append(o)
        start_of = [e - n for e, n in zip(ends, sizes)]
        bucket_end = off
        for i in range(len(sizes) - 1, -1, -1):
            acc += sizes[i]
            if acc >= elems or i == 0:
                cuts.append((start_of[i], bucket_end))
                bucket_end = start_of[i]
                acc = 0
        self.bounds = cuts               # cuts[0] is the LAST range of the buffer (first to be ready)
        self.tensor_bucket = [0] * len(sizes)
        for b, (s, e) in enumerate(self.bounds):
            for i in range(len(sizes)):
                if start_of[i] >= s and ends[i] <= e:
                    self.tensor_bucket[i] = b
        self.need = [0] * len(self.bounds)
        for b in self.tensor_bucket:
            self.need[b] += 1
        self.reset()

    def reset(self):
        self.pending = list(self.need)
        self.works = []
        self.launched = [False] * len(self.bounds)

    def mark_ready(self, tensor_index: int):
        b = self.tensor_bucket[tensor_index]
        self.pending[b] -= 1
        if self.pending[b] == 0 and not self.launched[b]:
            self._launch(b)

    def _launch(self, b: int):
        s, e = self.bounds[b]
        self.launched[b] = True
        self.works.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        for b in range(len(self.bounds)):
            if not self.launched[b]:
                self._launch(b)
        for w in self.works:
            w.wait()
        self.reset()


class DataParallel:
    """Wraps an iswm_b200 DeepLabV3 model + criterion for multi-rank training (see module docstring)."""

    def __init__(self, model, criterion, bucket_bytes: int = 25 << 20, group=None, metrics=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (backend 'nccl', one process per GPU)")
        self.model, self.criterion, self.group = model, criterion, group
        self.engine = model.engine()
        self.world = dist.get_world_size(group)
        self.bucket_bytes = bucket_bytes
        self.bucketer: Optional[GradBucketer] = None
        # identical replicas: rank 0's parameters AND buffers are canonical (DataParallel's replica 0)
        with torch.no_grad():
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, 0, group=group)
        criterion.hist_hook = self._allreduce_hist
        self.engine.grad_ready_hook = self._grad_ready
        self._index = {id(p): i for i, p in enumerate(model.parameters())}
        if metrics is not None:
            self.attach_metrics(metrics)

    def attach_metrics(self, metrics):
        """Validation under data parallelism (SURVEY 8e "Metric"): every rank accumulates the confusion counts of ITS shard on
        the device; `metrics.get_results()` / `.confusion_matrix` then all-reduce (SUM) the n*n int64 counters once, so every
        rank reports the whole validation set's numbers - bit-exact by construction (metrics/stream_metrics.py:122)."""
        metrics.process_group = self.group if self.group is not None else dist.group.WORLD
        return metrics

    def _allreduce_hist(self, hist: torch.Tensor):
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=self.group)

    def _grad_ready(self, p):
        if self.bucketer is None:
            sizes = [q.numel() for q in self.model.parameters()]
            self.bucketer = GradBucketer(self.engine.flat_g, sizes, self.bucket_bytes, self.group)
        self.bucketer.mark_ready(self._index[id(p)])

    def train_step(self, images, labels, optimizer=None):
        """forward -> global-batch weighted CE -> backward with overlapped bucketed all-reduce -> step.
        Returns the GLOBAL loss (sum over ranks of local numerators / global denominator)."""
        logits = self.model(images)
        loss = self.criterion(logits, labels)
        if optimizer is not None:
            optimizer.zero_grad()
        loss.backward()
        if self.bucketer is not None:
            self.bucketer.finish()
        gl = loss.detach().clone()
        dist.all_reduce(gl, op=dist.ReduceOp.SUM, group=self.group)
        if optimizer is not None:
            optimizer.step()
        return gl
