"""Time mtn_si_snr_pit_fwd (SI-SNR + PIT + SI-SNRi on device) against its HBM roofline (20 bytes per sample).

    python tools/score_bench.py [--batch 512] [--T 96000]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avse_challenge_b200 import scoring

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512); ap.add_argument("--T", type=int, default=96000)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(0)
src = torch.randn(a.batch, a.T, 2, device="cuda", generator=g) * 0.05
est = src * 0.9 + torch.randn(a.batch, a.T, 2, device="cuda", generator=g) * 0.005
mix = src.sum(-1)
for _ in range(3): scoring.si_snr_pit(est, src, mix)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters): m = scoring.si_snr_pit(est, src, mix)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
alg = a.batch * a.T * 20
print(json.dumps({"op": "si_snr_pit", "batch": a.batch, "T": a.T, "ms": round(ms, 4), "alg_GBps": round(alg / ms / 1e6, 1),
                  "frac_of_6541": round(alg / ms / 1e6 / 6541.1, 3), "mean_si_snr_db": float(m["si_snr"].mean())}))
