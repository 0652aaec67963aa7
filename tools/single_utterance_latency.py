"""Latency of separating ONE utterance (BASELINE config 1's shape: 4 s @ 8 kHz) on one GPU: the batch plan at B = 1
against the chunked-scan plan of the sequence-parallel driver run on a single rank (time cut into `sub_chunks` pieces:
summary pass, fold, seeded pass -- the scan's serial chain is `sub_chunks` times shorter, which is what bounds B = 1).

    python tools/single_utterance_latency.py [--hparams S] [--seconds 4] [--sub-chunks 4,8,16]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avse_challenge_b200 import CONFIGS, init_state_dicts, synth_mixture
from avse_challenge_b200.engine import SeparatorEngine
from avse_challenge_b200.parallel import SequenceParallelSeparator

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S"); ap.add_argument("--seconds", type=float, default=4.0)
ap.add_argument("--sample-rate", type=int, default=8000); ap.add_argument("--mode", default="fp32")
ap.add_argument("--sub-chunks", default="4,8,16"); ap.add_argument("--iters", type=int, default=30)
a = ap.parse_args()
hp = CONFIGS[a.hparams]
T = int(a.seconds * a.sample_rate) // 8 * 8
sds = init_state_dicts(hp, 1234)
mix = synth_mixture(1, T, a.sample_rate, seed=1234)[0].cuda()


def timed(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters


eng = SeparatorEngine(hp, sds, device="cuda", mode=a.mode, small_batch_plan=False)   # the batch plan, for comparison
ref = eng(mix)
res = {"shape": [a.hparams, 1, T, a.mode], "batch_plan_ms": round(timed(lambda: eng(mix)), 3), "chunked": []}
rms = ref.pow(2).mean().sqrt()
for sc in [int(x) for x in a.sub_chunks.split(",")]:
    eager = SequenceParallelSeparator(hp, sds, device="cuda", mode=a.mode, sub_chunks=sc, use_graph=False)
    graph = SequenceParallelSeparator(hp, sds, device="cuda", mode=a.mode, sub_chunks=sc, use_graph=True)
    est = graph(mix)
    res["chunked"].append({"sub_chunks": sc, "max_abs_diff/rms_vs_batch_plan": ((est - ref).abs().max() / rms).item(),
                           "eager_ms": round(timed(lambda: eager(mix)), 3), "graph_ms": round(timed(lambda: graph(mix)), 3)})
print(json.dumps(res), flush=True)
