"""Round-2 golden vectors from the REAL reference code (run in the build container; TEST INFRASTRUCTURE).

    python oracle/gen_golden_r2.py

  model_r50_os8_320.npz   reference deeplabv3plus_resnet50(num_classes=2, output_stride=8) in eval mode on a
                    1x3x320x320 input: the 40x40 feature map is larger than every ASPP rate (12 / 24 / 36,
                    network/modeling.py:27-33), so all 27 taps of the three dilated branches land INSIDE the image (the 56x72 /
                    48x40 fixtures only exercise the centre taps of the larger rates). Stored: the input seed + a checksum of
                    the regenerated input, the logits on the stride-4 lattice (the final x4 bilinear upsample,
                    network/utils.py:22, is smooth; the full-resolution map is checked through the oracle), plus 4096
                    full-resolution samples at fixed random positions.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from oracle.gen_golden import seeded_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_input(H=320, W=320, seed=21):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((1, 3, H, W), generator=g)


def sample_positions(H=320, W=320, n=4096, seed=22):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, H * W, (n,), generator=g)


def main():
    modeling, _ = ref_import.reference_modules()
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = modeling.deeplabv3plus_resnet50(num_classes=2, output_stride=8, pretrained_backbone=False)
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=1234))
    model.eval()
    x = golden_input()
    with torch.no_grad():
        logits = model(x)
    pos = sample_positions()
    out = {"x_seed": np.int64(21), "x_sum": np.float64(x.double().sum().item()), "x_head": x.flatten()[:16].numpy().copy(),
           "logits_s4": logits[:, :, ::4, ::4].numpy().copy(), "pos": pos.numpy(),
           "logits_at_pos": logits.flatten(2)[0][:, pos].numpy().copy(),
           "logits_absmax": np.float64(logits.abs().max().item()), "logits_norm": np.float64(logits.double().norm().item())}
    np.savez_compressed(os.path.join(OUT, "model_r50_os8_320.npz"), **out)
    print("wrote model_r50_os8_320.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
