"""Summarise one launch of an ncu --set full report: duration, pipe/dram numbers and the hottest SASS lines with stall reasons.
usage: python tools/ncu_stalls.py <report.ncu-rep> <launch index> [top N]"""
import csv, subprocess, sys, io
rep, k = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2]
ix = {h: i for i, h in enumerate(hdr)}
for w in ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"]:
    if w in ix:
        print(f"{w}: {data[ix[w]][:90]} {units[ix[w]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stallcols}
print("samples", tot, sorted(agg.items(), key=lambda kv: -kv[1])[:8])
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:topn]
for i in sorted(top):
    r = data[i]
    st = sorted([(h[6:], int(r[ix[h]] or 0)) for h in stallcols if int(r[ix[h]] or 0) > 0], key=lambda kv: -kv[1])[:2]
    print(i, r[ix["# Samples"]].rjust(5), r[ix["Instructions Executed"]].rjust(8), r[ix["Source"]][:64], st)
