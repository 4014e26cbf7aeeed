"""Chunks-per-tile sweep of the weight-gradient kernel (ISWM_WGRAD_G is read once per process): python tools/wgrad_g_sweep.py"""
import os
import subprocess
import sys

SHAPES = [  # B, Cin, H, W, Cout, k, tag
    (16, 64, 128, 128, 64, 3, "layer1.conv2"), (16, 128, 64, 64, 128, 3, "layer2.conv2"), (16, 304, 128, 128, 256, 3, "decoder.0"),
    (16, 256, 128, 128, 256, 3, "decoder.3"), (16, 256, 32, 32, 256, 3, "layer3.conv2"), (16, 512, 32, 32, 512, 3, "layer4.conv2"),
    (16, 2048, 32, 32, 256, 3, "aspp.1"),
    (16, 1024, 32, 32, 256, 1, "layer3.conv1"), (16, 256, 32, 32, 1024, 1, "layer3.conv3"), (16, 2048, 32, 32, 256, 1, "aspp.0"),
    (16, 2048, 32, 32, 512, 1, "layer4.conv1"), (16, 512, 32, 32, 2048, 1, "layer4.conv3"), (16, 1280, 32, 32, 256, 1, "aspp.project"),
    (16, 512, 64, 64, 128, 1, "layer2.conv1"), (16, 128, 64, 64, 512, 1, "layer2.conv3"), (16, 256, 128, 128, 64, 1, "layer1.conv1"),
    (16, 64, 128, 128, 256, 1, "layer1.conv3"), (16, 1024, 32, 32, 2048, 1, "layer4.0.ds"), (16, 512, 32, 32, 1024, 1, "layer3.0.ds"),
    (16, 256, 64, 64, 512, 1, "layer2.0.ds"), (16, 256, 128, 128, 48, 1, "low_proj"), (1, 160, 1, 1048576, 64, 1, "stem(im2col)"),
]

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from iswm_b200 import ops
    dev = "cuda:0"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for (B, Cin, H, W, Cout, k, tag) in SHAPES:
        x = torch.randn((B, H, W, Cin), device=dev).to(torch.bfloat16)
        dy = torch.randn((B, H, W, Cout), device=dev).to(torch.bfloat16)
        dw = torch.zeros((Cout, k * k, Cin), dtype=torch.float32, device=dev)
        d = ops.make_conv_desc(B, H, W, Cin, Cin, B, H, W, Cout, Cout, ops.conv_taps(k, 1))
        ts = []
        for _ in range(4):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.conv_wgrad(d, x, dy, dw); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"{tag} {min(ts[1:]):.1f}")
    sys.exit(0)

vals = [0, 1, 2, 3, 4, 5, 6]
table = {}
for v in vals:
    env = dict(os.environ)
    if v:
        env["ISWM_WGRAD_G"] = str(v)
    out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        tag, t = line.rsplit(" ", 1)
        table.setdefault(tag, {})[v] = float(t)
print("shape".ljust(16) + "".join(f"{('auto' if v == 0 else v):>8}" for v in vals))
for tag, row in table.items():
    print(tag.ljust(16) + "".join(f"{row.get(v, float('nan')):8.1f}" for v in vals))
