"""Per-call latency of `MambaBlocksSequential.forward(x, inference_params)` (the reference's decode interface) for a few
tokens per call: the one-launch kernel against the chunk kernels.   python tools/stack_step_latency.py [--hparams S] [--tokens 1]"""
import argparse, json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avse_challenge_b200 import CONFIGS, init_state_dicts, modules, stream_fused

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S"); ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--tokens", type=int, default=1); ap.add_argument("--steps", type=int, default=200)
a = ap.parse_args()
hp = CONFIGS[a.hparams].causal()
m = init_state_dicts(hp, 5)["masknet"]
net = modules.MambaBlocksSequential(hp.n_mamba, bidirectional=False, d_model=hp.d_model, fused_add_norm=False, rms_norm=True)
net.load_state_dict({k[len("mamba_net."):]: v for k, v in m.items() if k.startswith("mamba_net.")}, strict=True)
net.cuda()
x = torch.randn(a.batch, a.tokens, hp.d_model, device="cuda")
res = {}
for label, cap in (("one_launch", 32), ("chunk_kernels", 0)):
    stream_fused.MAX_FRAMES = cap
    ip = types.SimpleNamespace(seqlen_offset=0, key_value_memory_dict={})
    net(x, inference_params=ip); ip.seqlen_offset = a.tokens
    for _ in range(10): net(x, inference_params=ip)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps): net(x, inference_params=ip)
    e1.record(); torch.cuda.synchronize()
    res[label + "_ms_per_call"] = e0.elapsed_time(e1) / a.steps
print(json.dumps({"hparams": hp.name, "batch": a.batch, "tokens_per_call": a.tokens, **res}))
