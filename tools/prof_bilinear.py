import os, sys, torch
sys.path.insert(0, "/root/repo")
from iswm_b200 import _lib
L = _lib.lib()
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = lambda: torch.cuda.current_stream().cuda_stream
for (B, Hi, Wi, C, ld, tag) in [(16, 32, 32, 256, 304, "cfg2 train"), (8, 128, 128, 256, 304, "cfg4 predict")]:
    x = torch.randn((B, Hi, Wi, C), device=dev).to(torch.bfloat16)
    out = torch.empty((B, 4 * Hi, 4 * Wi, ld), dtype=torch.bfloat16, device=dev)
    for flag in ("0", "1"):
        os.environ["ISWM_BILINEAR_UP4"] = flag
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(L.iswm_bilinear_fwd(x.data_ptr(), C, B, Hi, Wi, C, 4 * Hi, 4 * Wi, out[..., ld - C:].data_ptr(), ld, st()), "b")
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        print(tag, "up4" if flag == "1" else "generic", f"{min(ts[1:]):.1f} us  {out.numel() / ld * C * 2 / min(ts[1:]) / 1e3:.0f} GB/s written")
