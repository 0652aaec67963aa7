"""Micro-benchmark of mtn_scan_fwd alone at a BASELINE shape (used for ncu captures and kernel iteration).

    python tools/scan_bench.py [--hparams S] [--batch 32] [--L 3999] [--mode fp32] [--iters 20]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avse_challenge_b200 import CONFIGS, ops, _lib
if os.environ.get("MTN_LIB"):   # dev knob of this tool only: time an experimental build of the library
    _lib.LIB_PATH = os.path.abspath(os.environ["MTN_LIB"])
    if os.environ["MTN_LIB"].endswith("_dev.so"):
        _lib.EXPECTED_ABI += 1000     # tools/devbuild.sh experiment builds identify themselves this way

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S"); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--L", type=int, default=3999); ap.add_argument("--mode", default="fp32")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--variants", default="0", help="comma list of MTN_SCAN_VARIANT values to time (first = reference for the diff)")
a = ap.parse_args()
hp = CONFIGS[a.hparams]; di, R = hp.d_inner, hp.dt_rank; nd = ops.n_dbl_for(R)
P = 2 if a.mode == "fp32" else 1
M = a.batch * a.L
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
u = (torch.randn(P, M, 2 * di, device=dev, generator=g) * (1.0 if P == 1 else 0.5)).to(torch.bfloat16)
dbl = torch.randn(M, 2 * nd, device=dev, generator=g) * 0.5
xz = torch.randn(M, 2 * di, device=dev, generator=g).to(torch.float32 if P == 2 else torch.bfloat16)
w_dt = torch.randn(2, di, R, device=dev, generator=g) * R ** -0.5
dt_bias = torch.randn(2, di, device=dev, generator=g) * 0.5 - 3.0
A2 = -torch.exp(torch.randn(2, di, 16, device=dev, generator=g) * 0.5 + 0.5) * ops.LOG2E
Dk = torch.randn(2, di, device=dev, generator=g)
y = torch.empty_like(u)
hout = torch.zeros(2, a.batch, di, 16, device=dev)
RP = ops.rp_for(R)
cols = torch.stack([dbl[:, :RP], dbl[:, nd:nd + RP]], dim=1)
hi = cols.to(torch.bfloat16); lo = (cols - hi.float()).to(torch.bfloat16)
dtp = torch.stack([hi, lo], dim=2).contiguous()            # [M, 2, 2, RP] as the x_proj epilogue writes it
run_full = lambda: ops.scan(u, dbl, xz, di, w_dt, dt_bias, A2, Dk, a.batch, a.L, di, R, y=y, dtp=dtp)
run_sum = lambda: ops.scan(u, dbl, xz, di, w_dt, dt_bias, A2, Dk, a.batch, a.L, di, R, h_out=hout, summary_only=True, dtp=dtp)
s_io = 4 if P == 2 else 2
alg = 2 * (M * (4 * di + 32) * s_io + (di * 16 + 2 * di) * 4)
y_ref = None
for var in a.variants.split(","):
    run = run_sum if var.startswith("S") else run_full  # "S" / "S<v>" = summary pass (no y) of variant 0 / <v>
    os.environ["MTN_SCAN_VARIANT"] = (var[1:] or "0") if var.startswith("S") else var
    y.zero_()
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    yv = y.float().sum(0)
    if y_ref is None:
        y_ref = yv.clone()
    diff = ((yv - y_ref).abs().max() / y_ref.pow(2).mean().sqrt()).item()
    print(json.dumps({"variant": var, "scan_ms": round(ms, 4), "alg_GBps": round(alg / ms / 1e6, 1),
                      "frac_of_6541": round(alg / ms / 1e6 / 6541.1, 4), "max_abs_diff_vs_first/rms": diff,
                      "y_checksum": [yv.double().sum().item(), yv.double().abs().sum().item()],
                      "mufu_bound_ms@1965": round(2 * M * di * 16 / (148 * 16 * 1.965e9) * 1e3, 4),
                      "shape": [a.hparams, a.batch, a.L, a.mode]}), flush=True)
