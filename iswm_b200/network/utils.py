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

    @torch.no_grad()
    def forward_lowres(self, x):
        """Inference only: the classifier's fp32 NHWC [B,H/4,W/4,num_classes] logits BEFORE the final bilinear
        upsample of network/utils.py:22 (eval-mode BatchNorm / Dropout regardless of .training). Consumers that fuse
        the upsample with what follows (iswm_b200.predict.predict_mask) read 1/16 of the pixels."""
        return self.engine().forward(x, train=False, lowres=True)
