"""CPU checks of SURVEY 8 row a1 / KAT-5: the drop-in constructors build the reference's module tree (state-dict keys and
shapes pinned by the reference golden fixtures, parameter counts, convolution inventory, OS8 / OS16 dilation plans) and
refuse to run without a GPU."""
import os

import numpy as np
import pytest
import torch

from iswm_b200.network import modeling


@pytest.mark.parametrize("name,ctor,os_", [("model_r50_os16.npz", modeling.deeplabv3plus_resnet50, 16),
                                           ("model_r101_os8.npz", modeling.deeplabv3plus_resnet101, 8)])
def test_state_dict_keys_and_shapes_match_reference(golden_dir, name, ctor, os_):
    g = np.load(os.path.join(golden_dir, name), allow_pickle=True)
    m = ctor(num_classes=2, output_stride=os_, pretrained_backbone=False)
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in g["shapes"]]
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"])


def test_kat5_structure_counts():
    r50 = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False)
    assert sum(p.numel() for p in r50.parameters()) == 40_347_298 and len(r50.state_dict()) == 374
    e = r50.engine()
    assert len(e.specs) == 63 and sum(1 for s in e.specs if s.bn is not None) == 62
    r101 = modeling._load_model("deeplabv3plus", "resnet101", 2, output_stride=8, pretrained_backbone=False)
    assert sum(p.numel() for p in r101.parameters()) == 59_339_426 and len(r101.engine().specs) == 114


@pytest.mark.parametrize("os_,rates,l3,l4", [(16, (6, 12, 18), (1, 1), (2, 1)), (8, (12, 24, 36), (2, 1), (4, 1))])
def test_output_stride_dilation_plan(os_, rates, l3, l4):
    """network/modeling.py:12-56: OS16 -> ASPP 6/12/18, replace_stride_with_dilation [F,F,T]; OS8 -> 12/24/36, [F,T,T]
    (layer3 / layer4 3x3 convolutions: (dilation, stride) of the blocks after the first; the first block of a dilated
    layer keeps the previous dilation, resnet.py:176-198)."""
    m = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=os_, pretrained_backbone=False)
    e = m.engine()
    assert tuple(s.dilation for s in e.aspp_branches[1:]) == rates
    by = {s.name: s for s in e.specs}
    assert (by["backbone.layer3.1.conv2"].dilation, by["backbone.layer3.1.conv2"].stride) == l3
    assert (by["backbone.layer4.1.conv2"].dilation, by["backbone.layer4.1.conv2"].stride) == l4
    assert by["backbone.layer2.0.conv2"].stride == 2 and by["backbone.conv1"].stride == 2
    first3 = by["backbone.layer3.0.conv2"]
    assert (first3.stride, first3.dilation) == ((2, 1) if os_ == 16 else (1, 1))
    first4 = by["backbone.layer4.0.conv2"]
    assert (first4.stride, first4.dilation) == ((1, 1) if os_ == 16 else (1, 2))


def test_unknown_backbone_raises_like_the_reference():
    with pytest.raises(NotImplementedError):                 # modeling.py:70-71
        modeling._load_model("deeplabv3plus", "mobilenetv2", 2, output_stride=16, pretrained_backbone=False)


def test_cpu_forward_is_refused():
    m = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=16, pretrained_backbone=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 64, 64))
    from iswm_b200.utils.loss import CrossEntropyLoss, FocalLoss
    with pytest.raises(RuntimeError):
        CrossEntropyLoss()(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError):
        FocalLoss(gamma=2)(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
