"""Parameter containers. They carry exactly the tensors nn.Conv2d / nn.BatchNorm2d would register
(same names, shapes, dtypes, init) so checkpoints written by the reference load unchanged, but they
compute nothing themselves: the engine (iswm_b200/engine.py) runs them with CUDA kernels."""
from __future__ import annotations

import torch
import torch.nn as nn


class ConvParams(nn.Module):
    """weight [Cout,Cin,k,k] fp32 (+bias) — the state of an nn.Conv2d."""

    def __init__(self, cin: int, cout: int, k: int, bias: bool = False):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = cin, cout, (k, k)
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        self.bias = nn.Parameter(torch.empty(cout)) if bias else None
        # nn.Conv2d.reset_parameters
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if bias:
            bound = 1.0 / (cin * k * k) ** 0.5
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, *_):
        raise RuntimeError("ConvParams holds parameters only; the model's forward runs through the CUDA engine")


class BNParams(nn.Module):
    """weight, bias, running_mean, running_var, num_batches_tracked — the state of an nn.BatchNorm2d."""

    def __init__(self, c: int):
        super().__init__()
        self.num_features = c
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def forward(self, *_):
        raise RuntimeError("BNParams holds parameters only; the model's forward runs through the CUDA engine")


class Slot(nn.Module):
    """Parameter-free placeholder keeping nn.Sequential indices aligned with the reference
    (ReLU / Dropout / AdaptiveAvgPool2d positions)."""

    def __init__(self, what: str = ""):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what

    def forward(self, *_):
        raise RuntimeError("placeholder module")
