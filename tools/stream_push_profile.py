"""Per-kernel CUDA-event times of one streaming push (eager launches; shows where the ~0.75 ms of a 20 ms chunk go).

    python tools/stream_push_profile.py [--hparams S] [--batch 1] [--chunk-ms 20] [--reps 50]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avse_challenge_b200 import CONFIGS, init_state_dicts
from avse_challenge_b200.engine import SeparatorEngine
from avse_challenge_b200.streaming import StreamingSeparator

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S"); ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--chunk-ms", type=float, default=20.0); ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--mode", default="fp32")
a = ap.parse_args()
hp = CONFIGS[a.hparams].causal()
n = max(16, int(round(a.chunk_ms * 1e-3 * hp.sample_rate)) // 8 * 8)
eng = SeparatorEngine(hp, init_state_dicts(hp, 1234), device="cuda", mode=a.mode, use_graph=False)
st = StreamingSeparator(eng, a.batch, use_graph=False)
x = torch.randn(a.batch, n, device="cuda") * 0.05
for _ in range(5):
    st.push(x)
agg = {}
for _ in range(a.reps):
    eng._prof = []
    st.push(x)
    torch.cuda.synchronize()
    for name, e0, e1 in eng._prof:
        v = agg.setdefault(name, [0, 0.0])
        v[0] += 1; v[1] += e0.elapsed_time(e1)
    eng._prof = None
out = {k: {"launches_per_push": c // a.reps, "us_per_launch": round(t / c * 1e3, 2), "us_per_push": round(t / a.reps * 1e3, 1)}
       for k, (c, t) in agg.items()}
print(json.dumps({"shape": [a.hparams, a.batch, n, a.mode], "per_op": out,
                  "sum_us_per_push": round(sum(v["us_per_push"] for v in out.values()), 1),
                  "note": "eager launches bracketed by events: each figure includes launch overhead, not just execution"}))
