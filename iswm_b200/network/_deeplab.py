"""DeepLabV3+ head parameter tree (reference: network/_deeplab.py:33-61 DeepLabHeadV3Plus,
:121-172 ASPP) and the top-level module whose forward runs on the CUDA engine."""
from __future__ import annotations

import torch
import torch.nn as nn

from .params import BNParams, ConvParams, Slot
from .utils import _SimpleSegmentationModel

__all__ = ["DeepLabV3", "DeepLabHeadV3Plus", "ASPP"]


class DeepLabV3(_SimpleSegmentationModel):
    """Same role and name as the reference's DeepLabV3 (network/_deeplab.py:16-31)."""


def _cbr(cin, cout, k):
    return nn.Sequential(ConvParams(cin, cout, k), BNParams(cout), Slot("ReLU"))


class ASPP(nn.Module):
    def __init__(self, in_channels, atrous_rates):
        super().__init__()
        self.rates = tuple(atrous_rates)
        branches = [_cbr(in_channels, 256, 1)] + [_cbr(in_channels, 256, 3) for _ in self.rates]
        branches.append(nn.Sequential(Slot("AdaptiveAvgPool2d(1)"), ConvParams(in_channels, 256, 1), BNParams(256), Slot("ReLU")))
        self.convs = nn.ModuleList(branches)
        self.project = nn.Sequential(ConvParams(5 * 256, 256, 1), BNParams(256), Slot("ReLU"), Slot("Dropout(0.1)"))


class DeepLabHeadV3Plus(nn.Module):
    def __init__(self, in_channels, low_level_channels, num_classes, aspp_dilate=(6, 12, 18)):
        super().__init__()
        self.project = _cbr(low_level_channels, 48, 1)
        self.aspp = ASPP(in_channels, aspp_dilate)
        self.classifier = nn.Sequential(
            ConvParams(304, 256, 3), BNParams(256), Slot("ReLU"),
            ConvParams(256, 256, 3), BNParams(256), Slot("ReLU"),     # second 3x3: ISWM's addition (_deeplab.py:48)
            ConvParams(256, num_classes, 1, bias=True))
        for m in self.modules():                                       # _deeplab.py:63-69
            if isinstance(m, ConvParams):
                nn.init.kaiming_normal_(m.weight)
