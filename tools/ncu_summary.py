#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into the tracked profiles/ directory.

  python tools/ncu_summary.py <tag> [--launches gpurun_out/<tag>_launches.csv] [--rep gpurun_out/<tag>_prof_*.ncu-rep ...]

Writes profiles/<tag>_launches.md (per-kernel launch count, device time, share of the step, DRAM bytes,
tensor-pipe activity) from the `--metrics ... --csv` launch list, profiles/<tag>_<rep>.md (selected raw-page
metrics per captured launch) from each `--set full` report, and refreshes profiles/traffic.json
(dram__bytes_read.sum + dram__bytes_write.sum per launch, per kernel) that bench.py reports as roofline.traffic.
ncu's per-launch times are cold-cache and serialised: compare SHARES, not absolutes.
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("iswm::", "")
    return re.sub(r"<.*", "", name)


def to_bytes(val: float, unit: str) -> float:
    u = unit.lower()
    return val * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(val: float, unit: str) -> float:
    u = unit.lower()
    return val * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(u, 1e-3)


def launches_summary(path: str, tag: str):
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    c = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "Grid Size", "Block Size")}
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= c["Metric Value"]:
            continue
        d = per.setdefault(r[c["ID"]], {"name": short(r[c["Kernel Name"]])})
        try:
            v = float(r[c["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        m, u = r[c["Metric Name"]], r[c["Metric Unit"]]
        if m == "gpu__time_duration.sum":
            d["us"] = to_us(v, u)
        elif m.startswith("dram__bytes"):
            d[m] = to_bytes(v, u)
        else:
            d[m] = v
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "tw": 0.0})
        a["n"] += 1
        a["us"] += d.get("us", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tw"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * d.get("us", 0.0)
    tot = sum(a["us"] for a in agg.values())
    out = [f"# {tag}: ncu launch list ({len(per)} launches, {tot / 1e3:.2f} ms of serialised device time)", "",
           f"source: `{os.path.relpath(path, ROOT)}` (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "sm__pipe_tensor_cycles_active... --clock-control none). Per-launch times are cold-cache and serialised: compare shares.", "",
           "| kernel | launches | total us | share | avg us | DRAM read MB | DRAM write MB | avg DRAM GB/s | tensor pipe active % (time-weighted) |",
           "|---|---|---|---|---|---|---|---|---|"]
    traffic = {}
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
        out.append(f"| {name} | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {a['us'] / a['n']:.1f} | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} | {gbs:.0f} | {a['tw'] / a['us'] if a['us'] else 0:.1f} |")
        traffic[name] = {"launches": a["n"], "dram_bytes_per_launch": (a["rd"] + a["wr"]) / a["n"], "avg_us": a["us"] / a["n"], "source": f"profiles/{tag}_launches.md"}
    os.makedirs(PROF, exist_ok=True)
    with open(os.path.join(PROF, f"{tag}_launches.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    tpath = os.path.join(PROF, "traffic.json")
    old = {}
    if os.path.exists(tpath):
        old = json.load(open(tpath))
    old.update(traffic)
    json.dump(old, open(tpath, "w"), indent=1, sort_keys=True)
    print("\n".join(out[:4 + 14]))


def rep_summary(path: str, tag: str):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print("empty report", path)
        return
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    keys = [k for k in RAW_KEYS if k in hdr]
    base = os.path.splitext(os.path.basename(path))[0]
    out = [f"# {base}: ncu --set full, raw-page extract", "", f"source: `{os.path.relpath(path, ROOT)}` (not tracked; regenerate with tools/gpu_round.sh)", "",
           "| launch | kernel | " + " | ".join(keys) + " |", "|---|---|" + "---|" * len(keys)]
    for n, r in enumerate(rows[2:]):
        vals = [f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip() for k in keys]
        out.append(f"| {n} | {short(r[name_i])} | " + " | ".join(vals) + " |")
    with open(os.path.join(PROF, f"{base}.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out[4:]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep", nargs="*", default=[])
    a = ap.parse_args()
    if a.launches:
        launches_summary(a.launches, a.tag)
    for r in a.rep:
        rep_summary(r, a.tag)


if __name__ == "__main__":
    main()
