"""Model constructors — same names, arguments and defaults as the reference
(network/modeling.py:12-83), plus `deeplabv3plus_resnet101` (what the reference reaches only through
`_load_model('deeplabv3plus', 'resnet101', ...)`)."""
from __future__ import annotations

from ._deeplab import DeepLabHeadV3Plus, DeepLabV3
from .backbone import resnet

__all__ = ["deeplabv3plus_resnet50", "deeplabv3plus_resnet101"]


def _segm_resnet(name, backbone_name, num_classes, output_stride, pretrained_backbone, in_channels=3):
    if output_stride == 8:                                   # modeling.py:14-19
        replace_stride_with_dilation = [False, True, True]
        aspp_dilate = [12, 24, 36]
    else:
        replace_stride_with_dilation = [False, False, True]
        aspp_dilate = [6, 12, 18]
    if backbone_name not in ("resnet50", "resnet101"):
        raise NotImplementedError(f"backbone {backbone_name!r}: only resnet50 / resnet101 are on the accelerated path")
    if in_channels != 3:
        raise NotImplementedError("in_channels != 3 is unreachable from the reference's public constructor (modeling.py:25-43)")
    if name != "deeplabv3plus":
        raise NotImplementedError(f"arch {name!r}: only deeplabv3plus is on the accelerated path")
    backbone = getattr(resnet, backbone_name)(pretrained=pretrained_backbone,
                                             replace_stride_with_dilation=replace_stride_with_dilation)
    classifier = DeepLabHeadV3Plus(2048, 256, num_classes, aspp_dilate)
    return DeepLabV3(backbone, classifier)


def _load_model(arch_type, backbone, num_classes, output_stride, pretrained_backbone, temporal=False,
                model_type="parallel", opts=None, in_channels=3):
    if backbone.startswith("resnet"):                        # modeling.py:59-71
        return _segm_resnet(arch_type, backbone, num_classes, output_stride=output_stride,
                            pretrained_backbone=pretrained_backbone, in_channels=in_channels)
    raise NotImplementedError


def deeplabv3plus_resnet50(num_classes=21, output_stride=8, pretrained_backbone=True):
    """DeepLabV3+ with a ResNet-50 backbone (reference: network/modeling.py:75-83)."""
    return _load_model("deeplabv3plus", "resnet50", num_classes, output_stride=output_stride,
                       pretrained_backbone=pretrained_backbone)


def deeplabv3plus_resnet101(num_classes=21, output_stride=8, pretrained_backbone=True):
    """DeepLabV3+ with a ResNet-101 backbone (the reference's `_load_model(..., 'resnet101', ...)`)."""
    return _load_model("deeplabv3plus", "resnet101", num_classes, output_stride=output_stride,
                       pretrained_backbone=pretrained_backbone)
