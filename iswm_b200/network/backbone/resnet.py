"""ResNet-50/101 parameter tree with torchvision/reference key names
(reference: network/backbone/resnet.py:78-120 Bottleneck, :124-198 ResNet._make_layer)."""
from __future__ import annotations

import torch.nn as nn

from ..params import BNParams, ConvParams, Slot

__all__ = ["ResNetTrunk", "resnet50", "resnet101"]


class BottleneckParams(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride, dilation, with_downsample):
        super().__init__()
        self.conv1 = ConvParams(inplanes, planes, 1)
        self.bn1 = BNParams(planes)
        self.conv2 = ConvParams(planes, planes, 3)
        self.bn2 = BNParams(planes)
        self.conv3 = ConvParams(planes, planes * 4, 1)
        self.bn3 = BNParams(planes * 4)
        self.downsample = nn.Sequential(ConvParams(inplanes, planes * 4, 1), BNParams(planes * 4)) if with_downsample else None
        self.stride, self.dilation = stride, dilation


class ResNetTrunk(nn.Module):
    """conv1/bn1/relu/maxpool/layer1..4 — what IntermediateLayerGetter keeps (network/utils.py:62-75)."""

    def __init__(self, blocks, replace_stride_with_dilation=(False, False, False)):
        super().__init__()
        if len(replace_stride_with_dilation) != 3:
            raise ValueError("replace_stride_with_dilation should be None or a 3-element tuple, got {}".format(replace_stride_with_dilation))
        self.conv1 = ConvParams(3, 64, 7)
        self.bn1 = BNParams(64)
        self.relu = Slot("ReLU")
        self.maxpool = Slot("MaxPool2d(3, 2, 1)")
        inplanes, dilation = 64, 1
        for li, (planes, n, stride) in enumerate(zip((64, 128, 256, 512), blocks, (1, 2, 2, 2))):
            prev = dilation
            if li > 0 and replace_stride_with_dilation[li - 1]:      # resnet.py:180-182
                dilation *= stride
                stride = 1
            mods = [BottleneckParams(inplanes, planes, stride, prev, stride != 1 or inplanes != planes * 4)]   # :190-191
            inplanes = planes * 4
            mods += [BottleneckParams(inplanes, planes, 1, dilation, False) for _ in range(1, n)]
            setattr(self, f"layer{li + 1}", nn.Sequential(*mods))
        for m in self.modules():                                       # resnet.py:158-163
            if isinstance(m, ConvParams):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


def _resnet(blocks, pretrained, replace_stride_with_dilation):
    if pretrained:
        # resnet.py:218-224 downloads ImageNet weights; keep the reference's behaviour of failing
        # loudly when that is impossible rather than silently training from scratch
        from torch.hub import load_state_dict_from_url
        trunk = ResNetTrunk(blocks, replace_stride_with_dilation)
        url = {(3, 4, 6, 3): "https://download.pytorch.org/models/resnet50-19c8e357.pth",
               (3, 4, 23, 3): "https://download.pytorch.org/models/resnet101-5d3b4d8f.pth"}[tuple(blocks)]
        sd = load_state_dict_from_url(url, progress=True)
        trunk.load_state_dict({k: v for k, v in sd.items() if not k.startswith("fc.")}, strict=True)
        return trunk
    return ResNetTrunk(blocks, replace_stride_with_dilation)


def resnet50(pretrained=False, progress=True, replace_stride_with_dilation=(False, False, False), **_):
    return _resnet((3, 4, 6, 3), pretrained, replace_stride_with_dilation)


def resnet101(pretrained=False, progress=True, replace_stride_with_dilation=(False, False, False), **_):
    return _resnet((3, 4, 23, 3), pretrained, replace_stride_with_dilation)
