"""Micro-benchmark of mtn_conv_silu alone (used for ncu captures).  python tools/conv_bench.py [--hparams L] [--batch 64] [--mode bf16]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avse_challenge_b200 import CONFIGS, ops, _lib
if os.environ.get("MTN_LIB"):   # dev knob of this tool only: time an experimental build of the library
    _lib.LIB_PATH = os.path.abspath(os.environ["MTN_LIB"])

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="L"); ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--L", type=int, default=3999); ap.add_argument("--mode", default="bf16"); ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--variants", default="0", help="comma list of MTN_CONV_VARIANT values (experiment builds only)")
a = ap.parse_args()
hp = CONFIGS[a.hparams]; di = hp.d_inner
P = 2 if a.mode == "fp32" else 1
M = a.batch * a.L
xz = torch.randn(M, 2 * di, device="cuda").to(torch.float32 if P == 2 else torch.bfloat16)
w = torch.randn(2, di, 4, device="cuda") * 0.5; b = torch.randn(2, di, device="cuda") * 0.5
u = torch.empty(P, M, 2 * di, dtype=torch.bfloat16, device="cuda")
run = lambda: ops.conv_silu(xz, w, b, a.batch, a.L, di, P, u=u)
u_ref = None
for var in a.variants.split(","):
    os.environ["MTN_CONV_VARIANT"] = var
    u.zero_()
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    if u_ref is None: u_ref = u.clone()
    s_in = 4 if P == 2 else 2
    byts = M * (di * s_in + 2 * di * 2 * P)
    print(json.dumps({"variant": var, "conv_ms": round(ms, 4), "GBps": round(byts / ms / 1e6, 1), "frac_of_6541": round(byts / ms / 1e6 / 6541.1, 3),
                      "identical_to_first": bool(torch.equal(u, u_ref)), "shape": [a.hparams, a.batch, a.L, a.mode]}), flush=True)
