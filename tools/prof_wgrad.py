"""Weight-gradient shapes of a cfg2 step, one launch each after a warm-up (for ncu): python tools/prof_wgrad.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iswm_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def wgrad(B, Cin, H, W, Cout, k, tag, dil=1):
    x = torch.randn((B, H, W, Cin), device=dev).to(torch.bfloat16)
    dy = torch.randn((B, H, W, Cout), device=dev).to(torch.bfloat16)
    dw = torch.zeros((Cout, k * k, Cin), dtype=torch.float32, device=dev)
    d = ops.make_conv_desc(B, H, W, Cin, Cin, B, H, W, Cout, Cout, ops.conv_taps(k, dil))
    fl = 2.0 * B * H * W * Cout * Cin * k * k
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv_wgrad(d, x, dy, dw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = min(ts[1:]) if len(ts) > 1 else ts[0]
    print(f"wgrad {tag:22s} B{B} {Cin}->{Cout} k{k} {H}x{W}: {t:7.1f} us  {fl / t / 1e6:7.1f} TF/s  in {(x.numel() + dy.numel()) * 2 / t / 1e3:7.1f} GB/s (abort={ops.abort_code()})", flush=True)


wgrad(16, 304, 128, 128, 256, 3, "decoder.0")
wgrad(16, 256, 128, 128, 256, 3, "decoder.3")
wgrad(16, 2048, 32, 32, 256, 1, "aspp.0")
wgrad(16, 2048, 32, 32, 256, 3, "aspp.1 (d6)", 6)
wgrad(16, 1280, 32, 32, 256, 1, "aspp.project")
wgrad(16, 64, 128, 128, 64, 3, "layer1.conv2")
wgrad(16, 256, 128, 128, 64, 1, "layer1.conv1")
wgrad(16, 64, 128, 128, 256, 1, "layer1.conv3")
wgrad(16, 128, 64, 64, 128, 3, "layer2.conv2")
wgrad(16, 1024, 32, 32, 256, 1, "layer3.conv1")
wgrad(16, 256, 32, 32, 1024, 1, "layer3.conv3")
wgrad(16, 256, 32, 32, 256, 3, "layer3.conv2")
wgrad(16, 512, 32, 32, 512, 3, "layer4.conv2", 2)
wgrad(16, 2048, 32, 32, 512, 1, "layer4.conv1")
wgrad(16, 1024, 32, 32, 2048, 1, "layer4.0.downsample")
