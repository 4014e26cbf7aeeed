"""Fused optimiser step for the hot path's tail (reference: train.py:421-431 builds
torch.optim.SGD(momentum=0.9, nesterov=True, weight_decay) at torch's default lr 1e-3 — `--lr` never
reaches it — and train.py:1049 steps it every iteration).

`FusedSGD` has torch.optim.SGD's update rule but runs as ONE kernel over the engine's flat fp32
master-weight / gradient / momentum buffers instead of a multi-tensor foreach sweep."""
from __future__ import annotations

import torch

from . import _lib


class FusedSGD:
    def __init__(self, model, lr: float = 1e-3, momentum: float = 0.9, weight_decay: float = 0.0, nesterov: bool = True):
        if nesterov and momentum <= 0:
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        self.model = getattr(model, "module", model)
        self.engine = self.model.engine()
        self.lr, self.momentum, self.weight_decay, self.nesterov = lr, momentum, weight_decay, nesterov
        self.param_groups = [{"lr": lr, "momentum": momentum, "weight_decay": weight_decay, "nesterov": nesterov}]
        self._mom = None
        self._steps = 0

    def zero_grad(self, set_to_none: bool = True):
        for p in self.engine._param_list():
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def step(self):
        eng = self.engine
        flat_w = eng.flatten_parameters()
        flat_g = eng.flat_g
        if flat_g is None:
            raise RuntimeError("FusedSGD.step() before any backward pass")
        if self._mom is None or self._mom.numel() != flat_w.numel() or self._mom.device != flat_w.device:
            self._mom = torch.zeros_like(flat_w)
            self._steps = 0
        lr = self.param_groups[0]["lr"]
        _lib.check(_lib.lib().iswm_sgd_step(flat_w.data_ptr(), flat_g.data_ptr(), self._mom.data_ptr(), flat_w.numel(),
                                            lr, self.momentum, self.weight_decay, 1 if self.nesterov else 0,
                                            1 if self._steps == 0 else 0, torch.cuda.current_stream().cuda_stream), "sgd_step")
        self._steps += 1
        eng.invalidate_packed()

    def state_dict(self):
        return {"momentum_buffer": self._mom, "steps": self._steps, "param_groups": self.param_groups}

    def load_state_dict(self, sd):
        self._mom, self._steps = sd["momentum_buffer"], sd["steps"]
        self.param_groups = sd["param_groups"]


class CosineAnnealingLR:
    """torch.optim.lr_scheduler.CosineAnnealingLR(T_max=total_itrs, eta_min=lr*0.01) as set up at
    train.py:446-452 and stepped every iteration (train.py:1103); closed form."""

    def __init__(self, optimizer: FusedSGD, T_max: int, eta_min: float = 0.0):
        import math
        self._math = math
        self.opt, self.T_max, self.eta_min = optimizer, T_max, eta_min
        self.base_lr = optimizer.param_groups[0]["lr"]
        self.last_epoch = 0

    def step(self):
        self.last_epoch += 1
        m = self._math
        lr = self.eta_min + (self.base_lr - self.eta_min) * (1 + m.cos(m.pi * self.last_epoch / self.T_max)) / 2
        self.opt.param_groups[0]["lr"] = lr

    def get_last_lr(self):
        return [self.opt.param_groups[0]["lr"]]
