"""Per-push time of the one-launch streaming push over (streams, channels per CTA): which cluster size to use when.
    python tools/stream_grid.py [--hparams S] [--frames 20]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avse_challenge_b200 import CONFIGS, init_state_dicts, _lib
from avse_challenge_b200.engine import SeparatorEngine
from avse_challenge_b200.streaming import StreamingSeparator

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S"); ap.add_argument("--frames", type=int, default=20); ap.add_argument("--steps", type=int, default=100)
a = ap.parse_args()
dev = torch.device("cuda", 0)
hp = CONFIGS[a.hparams].causal()
eng = SeparatorEngine(hp, init_state_dicts(hp, 1234), device=dev, mode="fp32", use_graph=False)
n = 8 * a.frames
for B in (1, 2, 4, 8, 12, 16, 24, 32, 64):
    row = {"streams": B}
    for dsl in (32, 64, 128):
        sm = int(_lib.load().mtn_stream_push_smem_bytes(a.frames, hp.d_model, dsl))
        if sm == 0 or sm > 227 * 1024:
            continue
        st = StreamingSeparator(eng, B, fused=True, channels_per_cta=dsl)
        x = (0.1 * torch.randn(B, n + 8)).to(dev)
        st.push(x)
        xs = [(0.1 * torch.randn(B, n)).to(dev) for _ in range(4)]
        for i in range(5): st.push(xs[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps): st.push(xs[i % 4])
        e1.record(); torch.cuda.synchronize()
        row[f"dsl{dsl}_ms"] = round(e0.elapsed_time(e1) / a.steps, 4)
    print(json.dumps(row), flush=True)
