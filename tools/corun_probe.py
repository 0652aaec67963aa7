"""Probe: can one kernel of the layer run on the SMs that a half-batch scan occupies (one scan CTA per SM)?

    python tools/corun_probe.py [--batch 16]

For each candidate kernel K (repeated so that it lasts about as long as the scan): time the scan alone, K alone, and both
launched together on two streams.  `overlap` = (t_scan + t_K - t_both) / min(t_scan, t_K): 1 = fully concurrent, 0 = serialised.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avse_challenge_b200 import CONFIGS, init_state_dicts, ops, _lib
from avse_challenge_b200.engine import SeparatorEngine, Workspace

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S"); ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--T", type=int, default=32000); ap.add_argument("--mode", default="fp32")
ap.add_argument("--share", default="1")
a = ap.parse_args()
dev = torch.device("cuda", 0)
hp = CONFIGS[a.hparams]
eng = SeparatorEngine(hp, init_state_dicts(hp, 1234), device=dev, mode=a.mode, small_batch_plan=False, use_graph=False)
wsA, wsB = Workspace(hp, a.batch, a.T, dev, a.mode), Workspace(hp, a.batch, a.T, dev, a.mode)
for ws in (wsA, wsB):
    ws.mix.normal_(0, 0.05)
    eng._run(ws)
torch.cuda.synchronize()
lw = eng.w.layers[0]
di, R, nd, D, P = hp.d_inner, hp.dt_rank, eng.n_dbl, hp.d_model, eng.P
M = wsA.M
scan = lambda: ops.scan(wsA.u, wsA.dbl, wsA.xz, di, lw["w_dt"], lw["dt_bias"], lw["A2"], lw["D"], wsA.batch, wsA.L, di, R, y=wsA.y, dir_mask=3)
cands = {
    "add_rmsnorm": lambda: ops.add_rmsnorm(wsB.h, wsB.res, True, lw["norm"], P, xn=wsB.xn),
    "gemm_in_proj": lambda: ops.gemm(wsB.xn, lw["w_in"], M, 2 * di, D, out=wsB.xz, epilogue=_lib.EPI_INPROJ, epi_param=di),
    "conv_silu": lambda: ops.conv_silu(wsB.xz, lw["conv_w"], lw["conv_b"], wsB.batch, wsB.L, di, P, u=wsB.u, dir_mask=3),
    "gemm_x_proj": lambda: ops.gemm(wsB.u, lw["w_x"], M, nd, di, out=wsB.dbl, groups=2, out_group_stride=nd),
    "gemm_out_proj": lambda: ops.gemm(wsB.y, lw["w_out"], M, D, 2 * di, out=wsB.h),
}
s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
main = torch.cuda.current_stream()


def run(fa, fb, reps_b, iters=5):
    best = None
    for _ in range(iters + 2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        ea = torch.cuda.Event(enable_timing=True); eb = torch.cuda.Event(enable_timing=True)
        e0.record(main); s1.wait_event(e0); s2.wait_event(e0)
        if fa is not None:
            with torch.cuda.stream(s1): fa()
        ea.record(s1)
        if fb is not None:
            with torch.cuda.stream(s2):
                for _ in range(reps_b): fb()
        eb.record(s2)
        main.wait_event(ea); main.wait_event(eb); e1.record(main)
        torch.cuda.synchronize()
        t = (e0.elapsed_time(e1), e0.elapsed_time(ea), e0.elapsed_time(eb))
        best = t if best is None or t[0] < best[0] else best
    return best


for share in [int(c) for c in a.share.split(",")]:
    os.environ["MTN_GEMM_SHARE"] = str(share)
    t_scan = run(scan, None, 0)[0]
    for name, fb in cands.items():
        t1 = run(None, fb, 1)[0]
        reps = max(1, int(round(0.8 * t_scan / t1)))
        tb = run(None, fb, reps)[0]
        both = run(scan, fb, reps)
        print(json.dumps({"share": share, "kernel": name, "reps": reps, "scan_ms": round(t_scan, 4), "k_ms": round(tb, 4),
                          "both_ms": round(both[0], 4), "scan_in_both": round(both[1], 4), "k_in_both": round(both[2], 4),
                          "overlap": round((t_scan + tb - both[0]) / min(t_scan, tb), 3)}), flush=True)
