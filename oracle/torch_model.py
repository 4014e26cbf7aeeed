"""Plain-PyTorch fp32 restatement of the reference DeepLabV3+ graph — TEST INFRASTRUCTURE ONLY.

This is the floating-point oracle for the network part of the hot path: a compact,
independent re-expression of what network/modeling.py:12-83, network/_deeplab.py:33-172,
network/utils.py:16-93 and network/backbone/resnet.py:78-198 compute, with the SAME
state_dict key names and shapes, so weights move freely between the reference, this oracle
and the CUDA implementation. It runs on CPU (or any torch device) through stock torch.nn
ops. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import it.

Pinning: oracle/gen_golden.py loads identical weights into the real reference modules
(imported from /root/reference in the build container) and into this restatement and
records reference outputs in tests/golden/model_r50_os16.npz; tests/test_oracle_golden.py
checks this file against those vectors on every run.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

_RESNET_BLOCKS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3)}


def _conv(cin, cout, k, stride=1, dilation=1, bias=False):
    return nn.Conv2d(cin, cout, k, stride=stride, padding=dilation * (k // 2), dilation=dilation, bias=bias)


class _Bottleneck(nn.Module):
    """resnet.py:78-120: 1x1 -> 3x3(stride, dilation) -> 1x1(x4), BN after each, residual, ReLU."""

    def __init__(self, inplanes, planes, stride, dilation, downsample):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 1)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = _conv(planes, planes, 3, stride, dilation)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = _conv(planes, planes * 4, 1)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        if downsample:
            self.downsample = nn.Sequential(_conv(inplanes, planes * 4, 1, stride), nn.BatchNorm2d(planes * 4))
        else:
            self.downsample = None

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return F.relu(y + idt)


def _make_layers(blocks, replace_stride_with_dilation):
    """resnet.py:149-155, :176-198 incl. the 'first block keeps the previous dilation' rule."""
    layers = OrderedDict()
    inplanes, dilation = 64, 1
    for li, (planes, n, stride) in enumerate(zip((64, 128, 256, 512), blocks, (1, 2, 2, 2))):
        dilate = li > 0 and replace_stride_with_dilation[li - 1]
        prev = dilation
        if dilate:
            dilation *= stride
            stride = 1
        mods = [_Bottleneck(inplanes, planes, stride, prev, stride != 1 or inplanes != planes * 4)]
        inplanes = planes * 4
        mods += [_Bottleneck(inplanes, planes, 1, dilation, False) for _ in range(1, n)]
        layers[f"layer{li + 1}"] = nn.Sequential(*mods)
    return layers


class _Backbone(nn.ModuleDict):
    """IntermediateLayerGetter over a ResNet truncated after layer4 (network/utils.py:62-93)."""

    def __init__(self, name, replace_stride_with_dilation):
        mods = OrderedDict()
        mods["conv1"] = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        mods["bn1"] = nn.BatchNorm2d(64)
        mods["relu"] = nn.ReLU(inplace=True)
        mods["maxpool"] = nn.MaxPool2d(3, 2, 1)
        mods.update(_make_layers(_RESNET_BLOCKS[name], replace_stride_with_dilation))
        super().__init__(mods)

    def forward(self, x):
        out = {}
        for name, m in self.items():
            x = m(x)
            if name == "layer1":
                out["low_level"] = x
        out["out"] = x
        return out


class _ASPPPool(nn.Sequential):
    """_deeplab.py:130-141 (index 0 is the parameter-free pool so the conv/BN keep keys 1 and 2)."""

    def __init__(self, cin, cout):
        super().__init__(nn.AdaptiveAvgPool2d(1), _conv(cin, cout, 1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

    def forward(self, x):
        size = x.shape[-2:]
        return F.interpolate(super().forward(x), size=size, mode="bilinear", align_corners=False)


class _ASPP(nn.Module):
    """_deeplab.py:143-172."""

    def __init__(self, cin, rates):
        super().__init__()
        def cbr(k, d):
            return nn.Sequential(_conv(cin, 256, k, 1, d), nn.BatchNorm2d(256), nn.ReLU(inplace=True))
        self.convs = nn.ModuleList([cbr(1, 1)] + [cbr(3, r) for r in rates] + [_ASPPPool(cin, 256)])
        self.project = nn.Sequential(_conv(5 * 256, 256, 1), nn.BatchNorm2d(256), nn.ReLU(inplace=True), nn.Dropout(0.1))

    def forward(self, x):
        return self.project(torch.cat([c(x) for c in self.convs], dim=1))


class _Head(nn.Module):
    """DeepLabHeadV3Plus, _deeplab.py:33-61 (two 3x3 convs, the second an ISWM addition :48)."""

    def __init__(self, num_classes, rates):
        super().__init__()
        self.project = nn.Sequential(_conv(256, 48, 1), nn.BatchNorm2d(48), nn.ReLU(inplace=True))
        self.aspp = _ASPP(2048, rates)
        self.classifier = nn.Sequential(
            _conv(304, 256, 3), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
            _conv(256, 256, 3), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
            nn.Conv2d(256, num_classes, 1))

    def forward(self, feats):
        low = self.project(feats["low_level"])
        a = F.interpolate(self.aspp(feats["out"]), size=low.shape[2:], mode="bilinear", align_corners=False)
        return self.classifier(torch.cat([low, a], dim=1))


class OracleDeepLabV3Plus(nn.Module):
    """network/utils.py:16-25 forward; modeling.py:12-56 structure choices."""

    def __init__(self, backbone="resnet50", num_classes=2, output_stride=16):
        super().__init__()
        if output_stride == 8:                      # modeling.py:14-19
            rsd, rates = (False, True, True), (12, 24, 36)
        else:
            rsd, rates = (False, False, True), (6, 12, 18)
        self.backbone = _Backbone(backbone, rsd)
        self.classifier = _Head(num_classes, rates)
        for m in self.backbone.modules():            # resnet.py:158-163
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        for m in self.classifier.modules():          # _deeplab.py:63-69
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)

    def forward(self, x):
        size = x.shape[-2:]
        y = self.classifier(self.backbone(x))
        return F.interpolate(y, size=size, mode="bilinear", align_corners=False)


def oracle_model(backbone="resnet50", num_classes=2, output_stride=16):
    return OracleDeepLabV3Plus(backbone, num_classes, output_stride)


def train_step(model, images, labels, weight=None):
    """train.py:1045-1048: logits -> weighted CE (train.py:454-459) -> zero_grad -> backward."""
    crit = nn.CrossEntropyLoss(weight=weight, ignore_index=255, reduction="mean")
    model.zero_grad(set_to_none=True)
    logits = model(images)
    loss = crit(logits, labels)
    loss.backward()
    return logits.detach(), loss.detach()


@torch.no_grad()
def predict_step(model, images, threshold=0.5):
    """predict.py:262-278: softmax, foreground probability > threshold."""
    prob = torch.softmax(model(images), dim=1)
    return (prob[:, 1] > threshold).long(), prob[:, 1]
