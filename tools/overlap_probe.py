"""Probe: does the batch plan gain from running batch PARTS on separate streams, staggered so that one part's scan
(latency-bound, 26 % of HBM, no tensor pipe) overlaps the other parts' GEMM / conv / norm kernels?

    python tools/overlap_probe.py [--hparams S] [--batch 32] [--seconds 4] [--parts 2] [--conc 1]

``--conc c``: at most c scans in flight (scan k waits for scan k - c in issue order); 0 = no cross-stream ordering.
Prints per-kernel times of the batch plan at B and B / parts, then graph-replayed forward times of each plan and the
max-abs difference of the outputs (the batch plan is batch-independent bit for bit, so 0.0 is expected).
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avse_challenge_b200 import CONFIGS
from avse_challenge_b200.engine import SeparatorEngine
from avse_challenge_b200 import init_state_dicts

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S"); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--seconds", type=float, default=4.0); ap.add_argument("--sr", type=int, default=8000)
ap.add_argument("--mode", default="fp32"); ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--parts", default="2"); ap.add_argument("--conc", default="0,1")
ap.add_argument("--no-prof", action="store_true")
ap.add_argument("--share", default="0,1"); ap.add_argument("--prio", default="0,1")
a = ap.parse_args()
dev = torch.device("cuda", 0)
hp = CONFIGS[a.hparams]
T = int(a.seconds * a.sr)
eng = SeparatorEngine(hp, init_state_dicts(hp, 1234), device=dev, mode=a.mode, small_batch_plan=False)
mix = torch.randn(a.batch, T, device=dev) * 0.05


def timed(fn, iters):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if not a.no_prof:
    for b in sorted({a.batch} | {a.batch // int(p) for p in a.parts.split(",")}):
        prof = eng.profile_ops(b, T, steps=3)
        print(json.dumps({"profile_batch": b, **{k: round(v["ms"], 4) for k, v in prof.items()}}), flush=True)

ws = eng.workspace(a.batch, T)
ws.mix[:, :T].copy_(mix)
ref = eng.forward(mix)
ms_batch = timed(lambda: eng.forward_into_workspace(a.batch, T), a.iters)
print(json.dumps({"plan": "batch", "ms": round(ms_batch, 3)}), flush=True)
if not a.no_prof:
    os.environ["MTN_GEMM_SHARE"] = "1"
    prof = eng.profile_ops(a.batch // 2, T, steps=3)
    print(json.dumps({"profile_batch": a.batch // 2, "share": 1, **{k: round(v["ms"], 4) for k, v in prof.items()}}), flush=True)
    os.environ["MTN_GEMM_SHARE"] = "0"


def run_parts(wss, streams, conc, hi=None):
    main = torch.cuda.current_stream()
    start = torch.cuda.Event(); start.record(main)
    scans = []          # events of the scans in issue order

    def op(name, fn, *args, **kw):
        if name != "scan":
            return fn(*args, **kw)
        cur = torch.cuda.current_stream()
        s = cur
        if hi is not None:      # the scan goes to a high-priority stream: its CTAs are placed before pending CTAs of other kernels
            s = hi[streams.index(cur)]
            ev = torch.cuda.Event(); ev.record(cur); s.wait_event(ev)
        if conc > 0 and len(scans) >= conc:
            s.wait_event(scans[-conc])
        with torch.cuda.stream(s):
            out = fn(*args, **kw)
        ev = torch.cuda.Event(); ev.record(s); scans.append(ev)
        if hi is not None:
            cur.wait_event(ev)
        return out

    eng._op = op
    try:
        hpp, w, P = eng.hp, eng.w, eng.w.P
        from avse_challenge_b200 import ops, _lib
        N, D = hpp.enc_dim, hpp.d_model
        for s in streams: s.wait_event(start)
        for wsp, s in zip(wss, streams):
            with torch.cuda.stream(s):
                op("encoder_cln", ops.encoder_cln, wsp.mix, w.w_enc, w.gamma, w.beta, P, mix_w=wsp.mix_w, yn=wsp.yn, T=wsp.T)
                op("gemm_bottleneck", ops.gemm, wsp.yn, w.w_bot, wsp.M, D, N, out=wsp.h)
        for i, lw in enumerate(w.layers):
            for wsp, s in zip(wss, streams):
                with torch.cuda.stream(s):
                    eng._layer(wsp, lw, first=(i == 0))
        for wsp, s in zip(wss, streams):
            with torch.cuda.stream(s):
                op("add_rmsnorm", ops.add_rmsnorm, wsp.h, wsp.res, True, w.norm_f, P, xn=wsp.xn, beta=w.norm_f_b)
                op("gemm_mask", ops.gemm, wsp.xn, w.w_mask, wsp.M, hpp.n_spk * N, D, out=wsp.sep, epilogue=_lib.EPI_MASK,
                   epi_param=N, aux=wsp.mix_w)
                op("decoder", ops.decoder, wsp.sep, w.w_dec, wsp.batch, wsp.T, wsp.L, N, hpp.n_spk, est=wsp.est, frames=wsp.frames)
            done = torch.cuda.Event(); done.record(s); main.wait_event(done)
    finally:
        del eng._op


from avse_challenge_b200.engine import Workspace
for parts in [int(p) for p in a.parts.split(",")]:
    bp = a.batch // parts
    wss = [Workspace(hp, bp, T, dev, a.mode) for _ in range(parts)]
    for k, wsp in enumerate(wss):
        wsp.mix[:, :T].copy_(mix[k * bp:(k + 1) * bp])
    streams = [torch.cuda.Stream(device=dev) for _ in range(parts)]
    his = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(parts)]
    for conc in [int(c) for c in a.conc.split(",")]:
      for share in [int(c) for c in a.share.split(",")]:
        for prio in [int(c) for c in a.prio.split(",")]:
            os.environ["MTN_GEMM_SHARE"] = str(share)
            hi = his if prio else None
            run_parts(wss, streams, conc, hi)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run_parts(wss, streams, conc, hi)
            ms = timed(g.replay, a.iters)
            got = torch.cat([wsp.est for wsp in wss], dim=0)
            print(json.dumps({"plan": "parts", "parts": parts, "conc": conc, "share": share, "prio": prio, "ms": round(ms, 3),
                              "vs_batch": round(ms / ms_batch, 4), "max_abs_diff": (got - ref).abs().max().item()}), flush=True)
            os.environ["MTN_GEMM_SHARE"] = "0"
