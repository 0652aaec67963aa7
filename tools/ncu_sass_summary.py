"""Summarise an ncu report's SASS page: per-region stall reasons, opcode mix, hottest instructions.

    python tools/ncu_sass_summary.py gpurun_out/x.ncu-rep [per_tile_divisor]
Regions are split at EXIT instructions (warp-specialised roles compile to separate straight-line bodies).
"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); h, v = r[0], r[2]
for k in ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
          "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "launch__registers_per_thread", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg"]:
    if k in h: print(f"{k:70s} {v[h.index(k)]} {r[1][h.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]
isrc, iall, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
st = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
def op(r):
    s = r[isrc].split(); return (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
bounds = [0] + [n + 1 for n, r in enumerate(data) if op(r) == "EXIT"] + [len(data)]
tot = sum(int(r[iall] or 0) for r in data)
for a, b in zip(bounds, bounds[1:]):
    sm = sum(int(r[iall] or 0) for r in data[a:b]); ex = sum(int(r[iex] or 0) for r in data[a:b])
    if sm < 0.02 * tot: continue
    d = {hdr[i][6:]: sum(int(r[i] or 0) for r in data[a:b]) for i in st}
    print(f"\n== region [{a},{b}) exec {ex} ({ex/div:.0f}/unit) samples {sm} ({100*sm/tot:.0f}%)")
    print("   stalls:", {k: f"{100*x/sm:.0f}%" for k, x in sorted(d.items(), key=lambda z: -z[1])[:9]})
    c = Counter()
    for rr in data[a:b]: c[op(rr)] += int(rr[iex] or 0)
    print("   mix/unit:", [(k, round(x / div, 1)) for k, x in c.most_common(24)])
    top = sorted(range(a, b), key=lambda n: -int(data[n][iall] or 0))[:12]
    for n in sorted(top):
        rr = data[n]; dd = {hdr[i][6:]: int(rr[i] or 0) for i in st if int(rr[i] or 0) > 0}
        print(f"   {n:5d} {rr[isrc][:58]:58s} ex {rr[iex]:>9s} smp {rr[iall]:>5s}", dict(sorted(dd.items(), key=lambda z: -z[1])[:4]))
