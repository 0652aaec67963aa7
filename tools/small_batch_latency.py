"""Latency of small batches: batch plan vs the chunked-scan plan run on up to SMALL_BATCH_STREAMS utterances at a time.

    python tools/small_batch_latency.py [--hparams S] [--seconds 4] [--batches 1,2,3,4,6,8,12,16]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avse_challenge_b200 import CONFIGS, init_state_dicts
from avse_challenge_b200.engine import SeparatorEngine

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S"); ap.add_argument("--seconds", type=float, default=4.0)
ap.add_argument("--sr", type=int, default=8000); ap.add_argument("--mode", default="fp32")
ap.add_argument("--batches", default="1,2,3,4,6,8,12,16"); ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--streams", default="1,4,8")
a = ap.parse_args()
dev = torch.device("cuda", 0)
hp = CONFIGS[a.hparams]
T = int(a.seconds * a.sr)
eng = SeparatorEngine(hp, init_state_dicts(hp, 1234), device=dev, mode=a.mode)


def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters


for B in [int(b) for b in a.batches.split(",")]:
    mix = torch.randn(B, T, device=dev) * 0.05
    eng.small_batch_plan = False
    ref = eng(mix)
    rec = {"batch": B, "batch_plan_ms": round(timed(lambda: eng(mix)), 3)}
    eng.small_batch_plan = True
    eng.SMALL_BATCH_MAX = 1 << 20
    for n in [int(s) for s in a.streams.split(",")]:
        eng.SMALL_BATCH_STREAMS = n
        got = eng(mix)
        rec[f"chunked_x{n}_ms"] = round(timed(lambda: eng(mix)), 3)
        rec[f"chunked_x{n}_vs_batch_plan_over_rms"] = ((got - ref).abs().max() / ref.pow(2).mean().sqrt()).item()
    print(json.dumps(rec), flush=True)
