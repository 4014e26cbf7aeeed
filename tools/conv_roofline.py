"""Per-launch roofline of the convolution kernels from a `bench.py --profile-detail` table: for every fwd / dgrad /
wgrad launch the larger of (algorithmic FLOPs / tensor peak) and (operand + result bytes / HBM peak) is the floor;
prints the launches ordered by time lost above it. usage: python tools/conv_roofline.py gpurun_out/<tag>_detail.txt"""
import sys

import torch

sys.path.insert(0, ".")
from iswm_b200.network import modeling  # noqa: E402

TF, GBS = 1382.1e12, 6549.4e9


def main(path, backbone="resnet50", os_=16):
    ctor = modeling.deeplabv3plus_resnet50 if backbone == "resnet50" else modeling.deeplabv3plus_resnet101
    with torch.device("meta"):
        m = ctor(num_classes=2, output_stride=os_, pretrained_backbone=False)
    specs = {s.name: s for s in m.engine().specs}
    rows, bn = [], []
    merged = {}                                    # several launches of one data gradient (parity phases) count as one
    order = []
    for line in open(path):
        p = line.split()
        kind, name, gf, us = p[0], p[1], float(p[-6]), float(p[-4])
        key = (kind, name)
        if kind.startswith("bn_") or key not in merged:
            if not kind.startswith("bn_"):
                merged[key] = len(order)
            order.append([kind, name, gf, us])
        else:
            order[merged[key]][2] += gf
            order[merged[key]][3] += us
    for kind, name, gf, us in order:
        if kind.startswith("bn_"):                 # HBM-bound entries: the third column is algorithmic MB
            bn.append((us - gf * 1e6 / GBS * 1e6, kind, name, us, gf, gf * 1e6 / (us * 1e-6) / 1e9))
            continue
        if name not in specs:                      # fused launches (the K-concatenated ASPP data gradient): tensor floor only
            t_tc = gf * 1e9 / TF * 1e6
            rows.append((us - t_tc, kind, name, us, t_tc, 0.0, "tensor"))
            continue
        s = specs[name]
        taps = 49 if s.is_stem else s.k * s.k
        M = gf * 1e9 / (2.0 * s.cout * s.cin * taps)
        Min = M * (s.stride ** 2)
        if kind == "fwd":
            by = 2 * (Min * s.cin + M * s.cout)
        elif kind == "dgrad":
            by = 2 * (M * s.cout + Min * s.cin)
        else:
            by = 2 * (Min * s.cin + M * s.cout) + 4 * s.cout * s.cin * taps
        t_tc, t_hbm = gf * 1e9 / TF * 1e6, by / GBS * 1e6
        floor = max(t_tc, t_hbm)
        rows.append((us - floor, kind, name, us, t_tc, t_hbm, "tensor" if t_tc >= t_hbm else "hbm"))
    rows.sort(reverse=True)
    tot = sum(r[3] for r in rows)
    fl = sum(max(r[4], r[5]) for r in rows)
    print(f"{len(rows)} launches, {tot:.0f} us measured, {fl:.0f} us at the per-launch floor")
    for kind in ("fwd", "dgrad", "wgrad"):
        a = sum(r[3] for r in rows if r[1] == kind)
        b = sum(max(r[4], r[5]) for r in rows if r[1] == kind)
        print(f"  {kind:6s} {a:8.0f} us vs floor {b:8.0f} us")
    print(f"{'lost us':>8s} {'kind':6s} {'conv':42s} {'us':>8s} {'tensor':>8s} {'hbm':>8s} bound")
    for r in rows[:int(sys.argv[2]) if len(sys.argv) > 2 else 50]:
        print(f"{r[0]:8.1f} {r[1]:6s} {r[2]:42s} {r[3]:8.1f} {r[4]:8.1f} {r[5]:8.1f} {r[6]}")


    if bn:
        bn.sort(reverse=True)
        for kind in sorted(set(b[1] for b in bn)):
            t = sum(b[3] for b in bn if b[1] == kind)
            mb = sum(b[4] for b in bn if b[1] == kind)
            print(f"  {kind:14s} {t:8.0f} us  {mb / 1e3:7.2f} GB  {mb * 1e6 / (t * 1e-6) / 1e9:7.0f} GB/s ({mb * 1e6 / (t * 1e-6) / GBS:.2f} of peak)")
        print(f"{'lost us':>8s} {'kind':14s} {'unit':42s} {'us':>8s} {'MB':>8s} {'GB/s':>8s}")
        for b in bn[:int(sys.argv[2]) if len(sys.argv) > 2 else 50]:
            print(f"{b[0]:8.1f} {b[1]:14s} {b[2]:42s} {b[3]:8.1f} {b[4]:8.1f} {b[5]:8.0f}")


if __name__ == "__main__":
    main(sys.argv[1])
