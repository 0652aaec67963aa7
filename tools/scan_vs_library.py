"""Same-box GPU comparator for the scan (SURVEY 8c "secondary cross-checks"): the cub-BlockScan selective-scan
design the reference runs through ``selective_scan_cuda.fwd`` (mamba-ssm), as shipped pre-compiled for sm_100 in this
image's vllm wheel (``vllm._custom_ops.selective_scan_fwd``), against ``mtn_scan_fwd`` on the same problem.

It is a *comparator*, not an oracle and not part of the product: LIBRARY code, timed the way the reference calls it
(channel-first [B, di, L] tensors, delta materialised by a dt_proj GEMM, one kernel launch per direction;
``selective_scan_interface.py:187,218-220``).  The flips of the backward direction (``bimamba.py:237,253``) are NOT
timed for the library arm, which favours it.

    python tools/scan_vs_library.py [--hparams S] [--batch 32] [--L 3999] [--mode fp32] [--iters 20]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from avse_challenge_b200 import CONFIGS, ops

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S"); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--L", type=int, default=3999); ap.add_argument("--mode", default="fp32")
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
hp = CONFIGS[a.hparams]; di, R = hp.d_inner, hp.dt_rank; nd = ops.n_dbl_for(R)
P = 2 if a.mode == "fp32" else 1
io = torch.float32 if P == 2 else torch.bfloat16
Bn, L = a.batch, a.L
M = Bn * L
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)

# ---- one problem, two layouts
u32 = torch.randn(M, 2 * di, device=dev, generator=g) * 0.5
if P == 2:
    hi = u32.to(torch.bfloat16); lo = (u32 - hi.float()).to(torch.bfloat16)
    u = torch.stack([hi, lo]).contiguous()
    u_val = hi.float() + lo.float()
else:
    u = u32.to(torch.bfloat16)[None].contiguous()
    u_val = u[0].float()
dbl = torch.randn(M, 2 * nd, device=dev, generator=g) * 0.5
z_raw = torch.randn(M, 2 * di, device=dev, generator=g).to(io)
xz = F.silu(z_raw.float()).to(io)                            # the in_proj epilogue hands the scan silu(z)
w_dt = torch.randn(2, di, R, device=dev, generator=g) * R ** -0.5
dt_bias = torch.randn(2, di, device=dev, generator=g) * 0.5 - 3.0
A = -torch.exp(torch.randn(2, di, 16, device=dev, generator=g) * 0.5 + 0.5)
A2 = (A * ops.LOG2E).contiguous()
Dk = torch.randn(2, di, device=dev, generator=g)
y = torch.empty_like(u)


def ours():
    ops.scan(u, dbl, xz, 0, w_dt, dt_bias, A2, Dk, Bn, L, di, R, y=y)


def cf(x2d, width):   # [M, width] token-major -> [B, width, L] channel-first contiguous (the reference's layout)
    return x2d.reshape(Bn, L, width).transpose(1, 2).contiguous()


lib_in = []
for d in range(2):
    c0 = d * nd
    lib_in.append(dict(
        u=cf(u_val[:, d * di:(d + 1) * di], di).to(io),
        dt=dbl[:, c0:c0 + R].contiguous(),                                   # [M, R], the dt_proj GEMM's operand
        B=cf(dbl[:, c0 + R:c0 + R + 16], 16).unsqueeze(1).to(io).contiguous(),
        C=cf(dbl[:, c0 + R + 16:c0 + R + 32], 16).unsqueeze(1).to(io).contiguous(),
        z=cf(z_raw[:, d * di:(d + 1) * di].float(), di).to(io),
        state=torch.zeros(Bn, di, 16, device=dev)))

from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn   # noqa: E402  (library comparator)


def library(keep=None):
    for d in range(2):
        t = lib_in[d]
        # dt_proj: delta = W_dt @ dt^T, channel-first like ssi.py:187 (bias + softplus are applied inside the kernel)
        delta = torch.matmul(w_dt[d], t["dt"].reshape(Bn, L, R).transpose(1, 2)).to(io)
        zbuf = t["z"].clone() if keep is not None else t["z"]   # the kernel writes its gated output in place over z
        out = selective_scan_fn(t["u"], t["state"], delta, A[d], t["B"], t["C"], Dk[d], zbuf, dt_bias[d],
                                delta_softplus=True)
        if keep is not None:
            keep.append(out)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters


# ---- numerics cross-check on the forward direction (the backward one differs only by the time order)
ours(); torch.cuda.synchronize()
kept = []
library(kept); torch.cuda.synchronize()
y0 = y.float().sum(0)[:, :di].reshape(Bn, L, di).transpose(1, 2)           # ours: 0.5 * gated output (bimamba.py:253)
ref0 = 0.5 * kept[0].float()
# fp32 mode only: in bf16 mode the library rounds delta, B and C to bf16 (its I/O dtype) while mtn_scan_fwd reads the fp32
# x_proj output, so the two arms do not compute the same function there and only the timing is comparable
rel = ((y0 - ref0).abs().max() / ref0.pow(2).mean().sqrt()).item() if P == 2 else None

ms_ours = timed(ours)
ms_lib = timed(library)
s_io = 4 if P == 2 else 2
alg = 2 * (M * (4 * di + 32) * s_io + (di * 16 + 2 * di) * 4)
print(json.dumps({"shape": [a.hparams, Bn, L, a.mode],
                  "mtn_scan_fwd_ms": round(ms_ours, 4), "library_dtproj_plus_2_scans_ms": round(ms_lib, 4),
                  "speedup": round(ms_lib / ms_ours, 3),
                  "alg_GBps_ours": round(alg / ms_ours / 1e6, 1), "alg_GBps_library": round(alg / ms_lib / 1e6, 1),
                  "fwd_dir_max_abs_diff/rms": rel,
                  "library": "vllm._custom_ops.selective_scan_fwd (sm_100 build of the mamba-ssm cub-BlockScan kernel) + torch.matmul dt_proj",
                  "note": "library arm excludes the two flips of the backward direction"}), flush=True)
