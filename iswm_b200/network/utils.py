"""Top-level segmentation module (reference: network/utils.py:8-25 _SimpleSegmentationModel).

forward(x: float32 [B,3,H,W]) -> float32 [B,num_classes,H,W], differentiable w.r.t. every
parameter; .train()/.eval() switch BatchNorm and Dropout behaviour. The computation is the CUDA
engine's (iswm_b200/engine.py); this class only bridges it to torch.autograd so that
`loss.backward()` (train.py:1048) and torch.optim (train.py:424-442) work unchanged."""
from __future__ import annotations

import torch
import torch.nn as nn


class _EngineFn(torch.autograd.Function):
    """One autograd node for the whole network. Parameter gradients are written by the engine
    straight into its flat buffer (exposed as each parameter's .grad), so backward returns None
    for them instead of 374 temporaries."""

    @staticmethod
    def forward(ctx, x, engine, anchor, *params):
        ctx.engine = engine
        ctx.n_params = len(params)
        out = engine.forward(x, train=True)
        ctx.generation = engine.generation
        return out

    @staticmethod
    def backward(ctx, dlogits):
        if ctx.generation != ctx.engine.generation:
            raise RuntimeError("backward() of a forward pass that a later train-mode forward superseded: the engine keeps "
                               "ONE pending tape (no gradient accumulation over several forwards, like the reference loop)")
        ctx.engine.backward(dlogits)
        return (None, None, None) + (None,) * ctx.n_params


class _EngineTailFn(torch.autograd.Function):
    """forward + criterion as ONE autograd node with the fused tail (csrc/tail_fused.cu): the full-resolution logits and their
    gradient never exist; backward hands the classifier-output gradient straight to the engine."""

    @staticmethod
    def forward(ctx, x, labels, engine, criterion, anchor, *params):
        from .. import ops
        ctx.engine, ctx.n_params = engine, len(params)
        lo = engine.forward(x, train=True, lowres=True, fused_tail=True)
        ctx.generation = engine.generation
        w = criterion.weight
        if w is not None and w.device != lo.device:
            w = w.to(lo.device)
        ev = engine._prof_begin()
        dlo_acc, hist, num = ops.tail_fwd(lo, labels, w, criterion.ignore_index)
        engine._prof_end(ev, "hbm:tail_fwd", float(labels.numel() * labels.element_size() + 2 * lo.numel() * 4), "tail_fwd")
        if criterion.hist_hook is not None:
            criterion.hist_hook(hist)                  # data parallel: SUM over ranks -> global-batch denominator
        criterion.last_hist = hist
        ctx.tail = (dlo_acc, hist, w, criterion.ignore_index)
        return ops.tail_loss(num, w, hist, criterion.ignore_index)

    @staticmethod
    def backward(ctx, g):
        from .. import ops
        eng = ctx.engine
        if ctx.generation != eng.generation:
            raise RuntimeError("backward() of a forward pass that a later train-mode forward superseded: the engine keeps "
                               "ONE pending tape (no gradient accumulation over several forwards, like the reference loop)")
        dlo_acc, hist, w, ignore = ctx.tail
        ctx.tail = None
        g = g.to(device=dlo_acc.device, dtype=torch.float32).contiguous()
        scratch = getattr(eng, "_tail_scratch", None)
        if scratch is None or scratch.device != dlo_acc.device:
            scratch = eng._tail_scratch = torch.zeros(8200, dtype=torch.uint8, device=dlo_acc.device)
        def fill(dlo, bias_grad):
            ev = eng._prof_begin()
            ops.tail_bwd(dlo_acc, w, hist, ignore, g, dlo, bias_grad, scratch)
            eng._prof_end(ev, "hbm:tail_bwd", float(dlo_acc.numel() * 4 + dlo.numel() * 2), "tail_bwd")

        eng.backward_tail(fill)
        return (None, None, None, None, None) + (None,) * ctx.n_params


class _SimpleSegmentationModel(nn.Module):
    def __init__(self, backbone, classifier):
        super().__init__()
        self.backbone = backbone
        self.classifier = classifier
        self._engine_obj = None

    def engine(self):
        if self._engine_obj is None:
            from ..engine import Engine
            self._engine_obj = Engine(self)
        return self._engine_obj

    def forward(self, x):
        eng = self.engine()
        params = eng._param_list()                      # cached walk of the (static) module tree
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if self.training:
            if needs_grad:
                return _EngineFn.apply(x, eng, params[0], *params)
            return eng.forward(x, train=True)          # BN batch statistics, no tape kept for backward
        with torch.no_grad():                          # eval: BatchNorm folded into the conv epilogues
            return eng.forward(x, train=False)

    def forward_loss(self, x, labels, criterion):
        """`criterion(model(x), labels)` for the training loop (train.py:1045-1046) with the tail fused: final x4 upsample, weighted
        cross entropy and their backward run as three small kernels over the LOW-RES classifier output and the labels; the
        full-resolution logits and their gradient are never written. Falls back to the plain composition whenever the fused
        kernels do not apply (eval mode, no gradient, another criterion, more than two classes, sizes not divisible by 4)."""
        from ..utils.loss import CrossEntropyLoss
        eng = self.engine()
        params = eng._param_list()
        fusable = (self.training and torch.is_grad_enabled() and any(p.requires_grad for p in params)
                   and type(criterion) is CrossEntropyLoss and not criterion.check_labels and eng.cls.cout == 2
                   and x.dim() == 4 and x.shape[2] % 4 == 0 and x.shape[3] % 4 == 0
                   and labels.dim() == 3 and labels.dtype in (torch.uint8, torch.int32, torch.int64))
        if not fusable:
            return criterion(self(x), labels)
        return _EngineTailFn.apply(x, labels, eng, criterion, params[0], *params)

    @torch.no_grad()
    def forward_lowres(self, x):
        """Inference only: the classifier's fp32 NHWC [B,H/4,W/4,num_classes] logits BEFORE the final bilinear
        upsample of network/utils.py:22 (eval-mode BatchNorm / Dropout regardless of .training). Consumers that fuse
        the upsample with what follows (iswm_b200.predict.predict_mask) read 1/16 of the pixels."""
        return self.engine().forward(x, train=False, lowres=True)
