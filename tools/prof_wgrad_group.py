"""All weight gradients of one ResNet layer at cfg2 size: one after the other at full width (what the step does) against
side-by-side launches on several streams, each planned for a share of the SMs (iswm_conv_wgrad_ex max_ctas).
python tools/prof_wgrad_group.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iswm_b200 import _lib, ops  # noqa: E402

L = _lib.lib()
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

LAYERS = {
    "layer3": [(16, 1024, 32, 32, 256, 1), (16, 256, 32, 32, 256, 3), (16, 256, 32, 32, 1024, 1)] * 6 + [(16, 512, 32, 32, 1024, 1)],
    "layer2": [(16, 512, 64, 64, 128, 1), (16, 128, 64, 64, 128, 3), (16, 128, 64, 64, 512, 1)] * 4 + [(16, 256, 64, 64, 512, 1)],
    "layer1": [(16, 256, 128, 128, 64, 1), (16, 64, 128, 128, 64, 3), (16, 64, 128, 128, 256, 1)] * 3 + [(16, 64, 128, 128, 256, 1)],
    "layer4": [(16, 2048, 32, 32, 512, 1), (16, 512, 32, 32, 512, 3), (16, 512, 32, 32, 2048, 1)] * 3 + [(16, 1024, 32, 32, 2048, 1)],
}


def build(shapes):
    jobs = []
    for (B, Cin, H, W, Cout, k) in shapes:
        x = torch.randn((B, H, W, Cin), device=dev).to(torch.bfloat16)
        dy = torch.randn((B, H, W, Cout), device=dev).to(torch.bfloat16)
        dw = torch.zeros((Cout, k * k, Cin), dtype=torch.float32, device=dev)
        d = ops.make_conv_desc(B, H, W, Cin, Cin, B, H, W, Cout, Cout, ops.conv_taps(k, 1))
        jobs.append((d, x, dy, dw))
    return jobs


def run(jobs, nstreams, streams):
    main = torch.cuda.current_stream()
    if nstreams <= 1:
        for (d, x, dy, dw) in jobs:
            _lib.check(L.iswm_conv_wgrad_ex(C.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), 0, main.cuda_stream), "w")
        return
    ev = main.record_event()
    budget = 148 // nstreams
    for i, (d, x, dy, dw) in enumerate(jobs):
        st = streams[i % nstreams]
        if i < nstreams:
            st.wait_event(ev)
        _lib.check(L.iswm_conv_wgrad_ex(C.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), budget, st.cuda_stream), "w")
    for st in streams[:nstreams]:
        main.wait_stream(st)


streams = [torch.cuda.Stream() for _ in range(8)]
for name, shapes in LAYERS.items():
    jobs = build(shapes)
    out = []
    for ns in (1, 4, "grouped"):
        ts = []
        for _ in range(4):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if ns == "grouped":
                ops.conv_wgrad_grouped(jobs)
            else:
                run(jobs, ns, streams)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        out.append(f"{ns} {min(ts[1:]):7.1f} us")
    print(f"{name} ({len(jobs)} weight gradients): " + "  ".join(out) + f"  (abort={ops.abort_code()})", flush=True)
