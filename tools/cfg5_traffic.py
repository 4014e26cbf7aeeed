"""DRAM bytes of the three cfg5 kernels from their own ncu --set full captures (gpurun_out/<tag>_lossmetric_<kernel>.ncu-rep)
into profiles/traffic.json under the keys bench.py reads ("cfg5:class_hist", "cfg5:wce_fwd_bwd", "cfg5:argmax_confusion").
usage: python tools/cfg5_traffic.py <tag>"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
names = {"class_hist": "cfg5:class_hist", "wce2": "cfg5:wce_fwd_bwd", "argmax_confusion": "cfg5:argmax_confusion"}
tp = os.path.join(ROOT, "profiles", "traffic.json")
traffic = json.load(open(tp))
lines = []
for k, key in names.items():
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_lossmetric_{k}.ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep)
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2]
    ix = {h: i for i, h in enumerate(hdr)}
    def val(m):
        v, u = float(data[ix[m]].replace(",", "")), units[ix[m]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "us": 1, "usecond": 1, "ns": 1e-3, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(u, 1)
    rd, wr, us = val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), val("gpu__time_duration.sum")
    traffic[key] = {"launches": 1, "dram_bytes_per_launch": rd + wr, "avg_us": us, "source": f"profiles/{tag}_lossmetric.md (ncu --set full of bench.py --workload lossmetric, kernel {data[ix['Kernel Name']][:60]})"}
    lines.append(f"| {key} | {data[ix['Kernel Name']][:70]} | {us:.1f} us | {rd / 1e6:.1f} MB read | {wr / 1e6:.1f} MB written | {(rd + wr) / us / 1e3:.0f} GB/s |")
json.dump(traffic, open(tp, "w"), indent=1, sort_keys=True)
with open(os.path.join(ROOT, "profiles", f"{tag}_lossmetric.md"), "w") as f:
    f.write(f"# {tag}: cfg5 loss / metric kernels, one ncu --set full capture each (16x2x1024x1024 fp32 logits, int64 labels)\n\n| key | kernel | time under ncu | DRAM read | DRAM write | DRAM rate |\n|---|---|---|---|---|---|\n" + "\n".join(lines) + "\n")
print("\n".join(lines))
