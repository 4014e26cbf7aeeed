"""Precision-matched oracle — TEST INFRASTRUCTURE ONLY.

The same DeepLabV3+ graph as oracle/torch_model.py (hence the reference's network/_deeplab.py,
network/backbone/resnet.py, network/utils.py — see the citations there), evaluated on the host CPU
with stock torch ops, but with a bf16 rounding at EXACTLY the points where the CUDA engine stores
bf16 (packed weights, conv outputs, activations, every activation gradient), and fp32 everywhere
the engine accumulates in fp32. Against this oracle the kernels must agree to fp32
summation-order noise (tests use <= 5e-3), which proves the fused kernels implement the stated
mixed-precision algorithm — forward AND backward — independently of how far bf16 itself drifts
from the fp32 reference (that drift is measured separately against oracle/torch_model.py).

Operates on an OracleDeepLabV3Plus instance (same parameters / state_dict as the reference).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def q(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).float()


class _QGrad(torch.autograd.Function):
    """identity forward; rounds the gradient to bf16 (an activation gradient stored by the engine)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return q(g)


class _QBoth(torch.autograd.Function):
    """bf16 storage of a forward value whose gradient is also stored in bf16."""

    @staticmethod
    def forward(ctx, x):
        return q(x)

    @staticmethod
    def backward(ctx, g):
        return q(g)


class _BNTrainQ(torch.autograd.Function):
    """conv output (fp32 accumulator) -> bf16 raw -> [fp32 stats OF THE STORED bf16 VALUES, as the conv epilogue
    sums its staged tile] -> normalise (+res) (+ReLU) -> bf16, with the engine's explicit backward
    (iswm_bn_bwd_reduce / iswm_bn_bwd_apply)."""

    @staticmethod
    def forward(ctx, y32, gamma, beta, residual, relu, running_mean, running_var):
        M = y32.numel() // y32.shape[1]
        yq = q(y32)
        mean = yq.sum((0, 2, 3)) / M
        var = ((yq * yq).sum((0, 2, 3)) / M - mean * mean).clamp_min(0)
        invstd = torch.rsqrt(var + BN_EPS)
        scale = gamma * invstd
        shift = beta - mean * scale
        z = yq * scale[None, :, None, None] + shift[None, :, None, None]
        if residual is not None:
            z = z + residual
        if relu:
            z = F.relu(z)
        out = q(z)
        if running_mean is not None:
            with torch.no_grad():
                unbiased = var * (M / (M - 1)) if M > 1 else var
                running_mean.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean)
                running_var.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * unbiased)
        ctx.save_for_backward(yq, out, gamma, mean, invstd)
        ctx.relu, ctx.has_res, ctx.M = relu, residual is not None, M
        return out

    @staticmethod
    def backward(ctx, dout):
        yq, out, gamma, mean, invstd = ctx.saved_tensors
        M = ctx.M
        dz = q(dout)
        if ctx.relu:
            dz = dz * (out > 0)
        xhat = (yq - mean[None, :, None, None]) * invstd[None, :, None, None]
        s1 = dz.sum((0, 2, 3))
        s2 = (dz * xhat).sum((0, 2, 3))
        k = (gamma * invstd)[None, :, None, None]
        dy = q(k * (dz - (s1 / M)[None, :, None, None] - xhat * (s2 / M)[None, :, None, None]))
        dres = q(dz) if ctx.has_res else None
        return dy, s2, s1, dres, None, None, None


def _wq(w):
    """bf16-rounded weight value with a straight-through gradient to the fp32 master weight."""
    return q(w.detach()) + (w - w.detach())


def _conv(x, conv):
    return F.conv2d(x, _wq(conv.weight), None, conv.stride, conv.padding, conv.dilation)


TRACE = None      # set to a list to record every unit's output in call order (tools/layer_diff.py)


def _unit(x, conv, bn, train, relu=True, residual=None):
    out = _unit_impl(x, conv, bn, train, relu, residual)
    if TRACE is not None:
        TRACE.append(out.detach())
    return out


def _unit_impl(x, conv, bn, train, relu=True, residual=None):
    y = _conv(x, conv)
    if train:
        with torch.no_grad():
            bn.num_batches_tracked += 1
        return _BNTrainQ.apply(y, bn.weight, bn.bias, residual, relu, bn.running_mean, bn.running_var)
    scale = bn.weight / torch.sqrt(bn.running_var + BN_EPS)
    shift = bn.bias - bn.running_mean * scale
    z = y * scale[None, :, None, None] + shift[None, :, None, None]
    if residual is not None:
        z = z + residual
    return q(F.relu(z) if relu else z)


def forward_q(model, x: torch.Tensor, train: bool) -> torch.Tensor:
    """model: oracle.torch_model.OracleDeepLabV3Plus; returns fp32 logits [B,C,H,W]."""
    bb, head = model.backbone, model.classifier
    size = x.shape[-2:]
    a = _unit(q(x), bb["conv1"], bb["bn1"], train)
    a = F.max_pool2d(a, 3, 2, 1)
    low = None
    for lname in ("layer1", "layer2", "layer3", "layer4"):
        for blk in bb[lname]:
            idt = a if blk.downsample is None else _unit(a, blk.downsample[0], blk.downsample[1], train, relu=False)
            y = _unit(a, blk.conv1, blk.bn1, train)
            y = _unit(y, blk.conv2, blk.bn2, train)
            a = _unit(y, blk.conv3, blk.bn3, train, relu=True, residual=idt)
        if lname == "layer1":
            low = a
    feat = a
    lowp = _unit(low, head.project[0], head.project[1], train)
    aspp = head.aspp
    outs = [_unit(feat, aspp.convs[i][0], aspp.convs[i][1], train) for i in range(4)]
    pooled = _QBoth.apply(feat.mean((2, 3), keepdim=True))
    pv = _unit(pooled, aspp.convs[4][1], aspp.convs[4][2], train)
    outs.append(pv.expand(-1, -1, feat.shape[2], feat.shape[3]))
    ao = _unit(torch.cat(outs, 1), aspp.project[0], aspp.project[1], train)      # Dropout(0.1) omitted: parity runs use p = 0
    up = _QBoth.apply(F.interpolate(ao, size=lowp.shape[2:], mode="bilinear", align_corners=False))
    y = torch.cat([lowp, up], 1)
    cl = head.classifier
    y = _unit(y, cl[0], cl[1], train)
    y = _unit(y, cl[3], cl[4], train)
    lo = F.conv2d(y, _wq(cl[6].weight), cl[6].bias)
    lo = _QGrad.apply(lo)
    return F.interpolate(lo, size=size, mode="bilinear", align_corners=False)


def train_step_q(model, images, labels, weight=None):
    crit = torch.nn.CrossEntropyLoss(weight=weight, ignore_index=255, reduction="mean")
    model.zero_grad(set_to_none=True)
    logits = forward_q(model, images, True)
    loss = crit(logits, labels)
    loss.backward()
    return logits.detach(), loss.detach()
