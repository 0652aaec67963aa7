"""Micro-benchmark of mtn_gemm_fwd at the shapes of one Mamba-TasNet layer (kernel iteration / ncu captures).

    python tools/gemm_bench.py [--hparams S] [--batch 32] [--L 3999] [--mode fp32] [--iters 20]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avse_challenge_b200 import CONFIGS, ops, _lib
if os.environ.get("MTN_LIB"):   # dev knob of this tool only: time an experimental build of the library
    _lib.LIB_PATH = os.path.abspath(os.environ["MTN_LIB"])

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S"); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--L", type=int, default=3999); ap.add_argument("--mode", default="fp32")
ap.add_argument("--iters", type=int, default=20); ap.add_argument("--only", default="")
ap.add_argument("--ab", action="store_true", help="time every case with MTN_GEMM_2CTA=0 (one CTA per tile) and =2 (CTA pairs wherever the shape allows) and compare the outputs")
ap.add_argument("--ab-stages", action="store_true", help="time every case with MTN_GEMM_HALF_STAGES=0 (64-deep stages) and =1 (32-deep stages where the dispatcher uses them) and compare the outputs")
ap.add_argument("--ab-env", default="", help="NAME=v0,v1,...: time every case with the library knob NAME set to each value and compare the outputs (e.g. MTN_GEMM_TMA_STORE=0,1)")
a = ap.parse_args()
hp = CONFIGS[a.hparams]; D, N, di, R = hp.d_model, hp.enc_dim, hp.d_inner, hp.dt_rank; nd = ops.n_dbl_for(R)
P = 2 if a.mode == "fp32" else 1
M = a.batch * a.L
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
planes = lambda rows, cols: (rnd(P, rows, cols) * 0.5).to(torch.bfloat16)
xz_dt = torch.float32 if P == 2 else torch.bfloat16
cases = {
    # name: (A planes, W planes, M, N, K, kwargs, bytes moved (algorithmic))
    "in_proj": (planes(M, D), planes(2 * di, D), M, 2 * di, D,
                dict(epilogue=_lib.EPI_INPROJ, epi_param=di, out_bf16=P == 1, out=torch.empty(M, 2 * di, device=dev, dtype=xz_dt)),
                M * D * 2 * P + M * 2 * di * (4 if P == 2 else 2)),
    "in_proj_plain_store": (planes(M, D), planes(2 * di, D), M, 2 * di, D,      # in_proj's shape without the SiLU epilogue (diagnostic)
                            dict(out=torch.empty(M, 2 * di, device=dev)), M * D * 2 * P + M * 2 * di * 4),
    "x_proj": (planes(M, 2 * di), planes(2 * nd, di), M, nd, di,
               dict(groups=2, out_group_stride=nd, out=torch.empty(M, 2 * nd, device=dev)),
               M * 2 * di * 2 * P + M * 2 * nd * 4),
    "out_proj": (planes(M, 2 * di), planes(D, 2 * di), M, D, 2 * di, dict(out=torch.empty(M, D, device=dev)),
                 M * 2 * di * 2 * P + M * D * 4),
    "out_proj_resadd": (planes(M, 2 * di), planes(D, 2 * di), M, D, 2 * di,        # the default plan's out_proj: res += result
                        dict(out=rnd(M, D), epilogue=_lib.EPI_RESADD, epi_param=1), M * 2 * di * 2 * P + 2 * M * D * 4),
    "bottleneck": (planes(M, N), planes(D, N), M, D, N, dict(out=torch.empty(M, D, device=dev)),
                   M * N * 2 * P + M * D * 4),
    "mask": (planes(M, D), planes(2 * N, D), M, 2 * N, D,
             dict(epilogue=_lib.EPI_MASK, epi_param=N, aux=rnd(M, N).abs(), out=torch.empty(M, 2 * N, device=dev)),
             M * D * 2 * P + M * N * 4 + M * 2 * N * 4),
}
for name, (A, W, m, n, k, kw, nbytes) in cases.items():
    if a.only and name not in a.only.split(","):
        continue
    run = lambda: ops.gemm(A, W, m, n, k, **kw)
    ref_out = None
    env_name, env_vals = (a.ab_env.split("=")[0], a.ab_env.split("=")[1].split(",")) if a.ab_env else ("", [])
    for pairs in (("0", "2") if a.ab else ("s0", "s1", "s2") if a.ab_stages else ["e" + v for v in env_vals] if env_vals else (os.environ.get("MTN_GEMM_2CTA", "1"),)):
        if pairs.startswith("e"):
            os.environ[env_name] = pairs[1:]                # read by the library at every call
        elif pairs.startswith("s"):
            os.environ["MTN_GEMM_HALF_STAGES"] = pairs[1]   # read by the library at every call
        else:
            os.environ["MTN_GEMM_2CTA"] = pairs          # read by the library at every call
        kw["out"].zero_()
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        groups = kw.get("groups", 1)
        flops = 2.0 * m * n * k * groups * (3 if P == 2 else 1)
        rec = {"gemm": name, "cta_pairs": pairs, "M": m, "N": n, "K": k, "groups": groups, "planes": P, "ms": round(ms, 4),
               "GBps_algorithmic": round(nbytes / ms / 1e6, 1), "frac_hbm_6541": round(nbytes / ms / 1e6 / 6541.1, 3),
               "TFLOPs_issued": round(flops / ms / 1e9, 1), "frac_bf16_sustained_1372": round(flops / ms / 1e9 / 1371.8, 3)}
        if ref_out is None:
            ref_out = kw["out"].float().clone()
        else:
            rec["max_abs_diff_vs_one_cta"] = (kw["out"].float() - ref_out).abs().max().item()
        print(json.dumps(rec), flush=True)
