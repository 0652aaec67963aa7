"""conv + x_proj: the fused kernel (mtn_conv_xproj_fwd) against the two-kernel plan, one layer's shapes.
    python tools/convx_bench.py [--hparams S] [--batch 32] [--L 3999] [--mode fp32]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avse_challenge_b200 import CONFIGS, ops

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S"); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--L", type=int, default=3999); ap.add_argument("--mode", default="fp32"); ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
hp = CONFIGS[a.hparams]; di, R = hp.d_inner, hp.dt_rank; nd = ops.n_dbl_for(R); P = 2 if a.mode == "fp32" else 1
dev = "cuda"; M = a.batch * a.L
g = torch.Generator(device=dev).manual_seed(0)
xz = torch.randn(M, 2 * di, device=dev, generator=g)
if P == 1: xz = xz.to(torch.bfloat16)
conv_w = torch.randn(2, di, 4, device=dev, generator=g) * 0.5
conv_b = torch.randn(2, di, device=dev, generator=g) * 0.1
w_x = ops.split_planes(torch.randn(2 * nd, di, device=dev, generator=g) / di ** 0.5, P)
u = torch.empty(P, M, 2 * di, dtype=torch.bfloat16, device=dev); dbl = torch.empty(M, 2 * nd, device=dev)
u2 = torch.empty_like(u); dbl2 = torch.empty_like(dbl)

def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters

def two():
    ops.conv_silu(xz, conv_w, conv_b, a.batch, a.L, di, P, u=u)
    ops.gemm(u, w_x, M, nd, di, out=dbl, groups=2, out_group_stride=nd)
ms_conv = timed(lambda: ops.conv_silu(xz, conv_w, conv_b, a.batch, a.L, di, P, u=u))
ms_gemm = timed(lambda: ops.gemm(u, w_x, M, nd, di, out=dbl, groups=2, out_group_stride=nd))
ms_two = timed(two)
ms_fused = timed(lambda: ops.conv_xproj(xz, conv_w, conv_b, w_x, a.batch, a.L, di, P, nd, u=u2, dbl=dbl2))
xb = xz.element_size() * di
bytes_fused = M * (xb + P * 2 * di * 2 + 2 * nd * 4)
print(json.dumps({"shape": [a.hparams, a.batch, a.L, a.mode], "conv_ms": round(ms_conv, 4), "x_proj_ms": round(ms_gemm, 4),
                  "two_kernel_ms": round(ms_two, 4), "fused_ms": round(ms_fused, 4), "fused_GBps": round(bytes_fused / ms_fused / 1e6, 1),
                  "frac_hbm_6541": round(bytes_fused / ms_fused / 1e6 / 6541.1, 3), "bit_identical": bool(torch.equal(u, u2) and torch.equal(dbl, dbl2))}))
