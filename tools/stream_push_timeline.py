"""Phase timeline of one fused streaming push (mtn_stream_push_fwd): globaltimer stamps of stream 0 / cluster rank 0 at the
phase boundaries of every layer, plus the CUDA-event time of back-to-back pushes.
    python tools/stream_push_timeline.py [--hparams S] [--batch 1] [--frames 20]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avse_challenge_b200 import CONFIGS, init_state_dicts
from avse_challenge_b200.engine import SeparatorEngine
from avse_challenge_b200.streaming import StreamingSeparator

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--frames", type=int, default=20)
ap.add_argument("--steps", type=int, default=200)
a = ap.parse_args()
dev = torch.device("cuda", 0)
hp = CONFIGS[a.hparams].causal()
eng = SeparatorEngine(hp, init_state_dicts(hp, 1234), device=dev, mode="fp32", use_graph=False)
st = StreamingSeparator(eng, a.batch, fused=True)
n = 8 * a.frames
g = torch.Generator().manual_seed(0)
chunks = [(0.1 * torch.randn(a.batch, n, generator=g)).to(dev) for _ in range(8)]
st.push(torch.cat([chunks[0], chunks[1][:, :8]], dim=1).contiguous())
for i in range(10):
    st.push(chunks[i % 8])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps):
    st.push(chunks[i % 8])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
tl = torch.zeros((hp.n_mamba + 2) * 16, dtype=torch.int64, device=dev)
f = st._fused
f.run(chunks[0], st.in_tail, False, st._halo, st._h, st.state["ola_tail"], timeline=tl)
torch.cuda.synchronize()
t = tl.view(hp.n_mamba + 2, 16).cpu()
names = ["", "rmsnorm", "in_proj mma", "in_proj epi", "conv", "x_proj+store", "barrier B", "dbl sum", "dt_proj", "scan",
         "out_proj+store", "barrier C", "h sum+store", "barrier A"]
rows = {}
for k in range(1, 14):
    prev = t[1:1 + hp.n_mamba, k - 1].clone()
    prev[1:] = torch.where(torch.tensor(k == 1), t[1:hp.n_mamba, 13], prev[1:]) if k == 1 else prev[1:]
    if k == 1:
        prev[0] = t[0, 0]
    d = (t[1:1 + hp.n_mamba, k] - prev).float()
    rows[names[k]] = {"mean_ns": round(d.mean().item()), "min_ns": int(d.min().item()), "max_ns": int(d.max().item())}
head_ns = int(t[0, 0].item())
out = {"hparams": hp.name, "batch": a.batch, "frames": a.frames, "ms_per_push_events": ms,
       "layers_total_us": (t[hp.n_mamba, 13] - t[0, 0]).item() / 1e3,
       "tail_us": (t[hp.n_mamba + 1, 2] - t[hp.n_mamba + 1, 0]).item() / 1e3, "per_layer_phase": rows}
print(json.dumps(out))
