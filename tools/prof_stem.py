"""Stem 7x7/s2 at cfg2 size: row-tap form (kpitch 24 / 32 / 64) against the im2col form: fill + forward conv (+stats) + wgrad."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iswm_b200 import _lib, ops  # noqa: E402
from iswm_b200._lib import check  # noqa: E402

L = _lib.lib()
dev = "cuda:0"
B, H, W = 16, 512, 512
H1, W1 = H // 2, W // 2
img = torch.randn((B, 3, H, W), device=dev)
w = torch.randn((64, 3, 7, 7), device=dev) * 0.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = lambda: torch.cuda.current_stream().cuda_stream


def timed(fn, reps=4):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts[1:])


out = torch.empty((B, H1, W1, 64), dtype=torch.bfloat16, device=dev)
dy = torch.randn((B, H1, W1, 64), device=dev).to(torch.bfloat16)
stats = torch.zeros(128, dtype=torch.float64, device=dev)
taps = [((r - 3 - ((r + 1) & 1)) // 2, 0, (r + 1) & 1) for r in range(7)]
wp = torch.empty(64 * 448, dtype=torch.bfloat16, device=dev)
arr, nblk = _lib.fill_pack_jobs([(w.data_ptr(), wp.data_ptr(), 64, 3, 49, 64, 448, 2)])
jobs = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone().to(dev)
check(L.iswm_pack_weights_batched(jobs.data_ptr(), 1, nblk, st()))
for kp in (24, 32, 64):
    rows = torch.empty((2 * B, H1, W1, kp), dtype=torch.bfloat16, device=dev)
    t_fill = timed(lambda: check(L.iswm_stem_rows(img.data_ptr(), B, 3, H, W, H1, W1, kp, rows.data_ptr(), st())))
    d = ops.make_conv_desc(B, H1, W1, kp, kp, 2 * B, H1, W1, 64, 64, taps, flags=_lib.EPI_STATS)
    t_fwd = timed(lambda: ops.conv_igemm(d, rows, wp, out, stats=stats))
    d0 = ops.make_conv_desc(B, H1, W1, kp, kp, 2 * B, H1, W1, 64, 64, taps)
    t_fwd0 = timed(lambda: ops.conv_igemm(d0, rows, wp, out))
    acc = torch.zeros((64, 7, kp), dtype=torch.float32, device=dev)
    t_wg = timed(lambda: ops.conv_wgrad(d0, rows, dy, acc))
    print(f"rows kpitch {kp}: fill {t_fill:.1f} us  fwd+stats {t_fwd:.1f} us  fwd {t_fwd0:.1f} us  wgrad {t_wg:.1f} us  (abort={ops.abort_code()})", flush=True)
col = torch.empty((B * H1 * W1, 160), dtype=torch.bfloat16, device=dev)
t_fill = timed(lambda: check(L.iswm_stem_im2col(img.data_ptr(), B, 3, H, W, H1, W1, 160, col.data_ptr(), st())))
wpi = ops.pack_weight_fwd(w, stem=True)
M = B * H1 * W1
d = ops.make_conv_desc(1, 1, M, 160, 160, 1, 1, M, 64, 64, [(0, 0, 0)], flags=_lib.EPI_STATS)
t_fwd = timed(lambda: ops.conv_igemm(d, col, wpi, out, stats=stats))
d0 = ops.make_conv_desc(1, 1, M, 160, 160, 1, 1, M, 64, 64, [(0, 0, 0)])
acc = torch.zeros((64, 160), dtype=torch.float32, device=dev)
t_wg = timed(lambda: ops.conv_wgrad(d0, col, dy, acc))
print(f"im2col K=160: fill {t_fill:.1f} us  fwd+stats {t_fwd:.1f} us  wgrad {t_wg:.1f} us")
