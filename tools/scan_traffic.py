"""Measure the DRAM traffic of ONE scan launch with ncu and record it for bench.py's `roofline.traffic`.

    python tools/scan_traffic.py [--out gpurun_out/scan_traffic.json]      (on a GPU box; needs ncu)

For each bench shape (BASELINE config 2: S, 32 x 3999 frames, fp32 mode; config 3 per GPU at 1 / 8 GPUs: L, 256 / 32 x 3999,
bf16 mode) it runs `tools/scan_bench.py` under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` on the scan kernel and
writes {source_hash, launches: {key: {dram_bytes_per_launch, read, write, source}}}.  The hash covers the scan sources
(bench.scan_source_hash); bench.py prints `traffic: null` whenever the committed file was captured for other sources, so a
stale figure can never be quoted.  Copy the result to profiles/scan_traffic.json and commit it with the kernel change.
"""
import argparse, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import scan_source_hash  # noqa: E402

SHAPES = [("S", 32, 3999, "fp32"), ("L", 256, 3999, "bf16"), ("L", 32, 3999, "bf16")]

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "scan_traffic.json"))
a = ap.parse_args()
res = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of ONE mtn::scan_kernel_pair launch (both directions), "
                   "ncu on B200 via tools/scan_traffic.py; valid only for the scan sources with this hash",
       "source_hash": scan_source_hash(), "launches": {}}
for hp, b, L, mode in SHAPES:
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k",
           "regex:scan_kernel_pair", "-s", "3", "-c", "1", "--csv", sys.executable, os.path.join(ROOT, "tools", "scan_bench.py"),
           "--hparams", hp, "--batch", str(b), "--L", str(L), "--mode", mode, "--iters", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(out)) if len(r) > 6 and r[-3].startswith("dram__bytes")]
    vals = {}
    for r in rows:
        unit, v = r[-2], float(r[-1].replace(",", ""))
        vals[r[-3]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    if len(vals) != 2:
        print(f"{hp} b{b} {mode}: could not parse ncu output\n{out[-2000:]}", file=sys.stderr)
        continue
    rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
    res["launches"][f"{hp}_b{b}_L{L}_{mode}"] = {
        "dram_bytes_per_launch": int(rd + wr), "read": int(rd), "write": int(wr),
        "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, tools/scan_bench.py --hparams {hp} --batch {b} --L {L} --mode {mode}"}
    print(hp, b, L, mode, int(rd + wr), flush=True)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
print("wrote", a.out)
