"""Multi-GPU check + timing of the sequence-parallel long-form mode (run under torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/seqpar_check.py [--hparams S] [--seconds 20] [--sample-rate 16000] [--sub-chunks 32] [--exchange allgather]

Rank 0 also runs the unchunked single-GPU engine on the same recording and prints the difference."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from avse_challenge_b200 import CONFIGS, init_state_dicts, synth_mixture
from avse_challenge_b200.engine import SeparatorEngine
from avse_challenge_b200.parallel import SequenceParallelSeparator

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S"); ap.add_argument("--seconds", type=float, default=20.0)
ap.add_argument("--sample-rate", type=int, default=16000); ap.add_argument("--sub-chunks", type=int, default=32)
ap.add_argument("--exchange", default="allgather"); ap.add_argument("--mode", default="fp32")
ap.add_argument("--iters", type=int, default=3); ap.add_argument("--no-check", action="store_true")
a = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
hp = CONFIGS[a.hparams]
sds = init_state_dicts(hp, 1234)
T = int(a.seconds * a.sample_rate) // 8 * 8
mix, _ = synth_mixture(1, T, a.sample_rate, seed=1234)
sp = SequenceParallelSeparator(hp, sds, device=dev, mode=a.mode, sub_chunks=a.sub_chunks, exchange=a.exchange)
mix_d = mix.to(dev)
est = sp(mix_d)
torch.cuda.synchronize()
times = []
for _ in range(a.iters):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    est = sp(mix_d)
    torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
t = torch.tensor([min(times)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
out = {"world": world, "hparams": a.hparams, "T": T, "frames": hp.frames(T), "sub_chunks": a.sub_chunks,
       "exchange": a.exchange, "mode": a.mode, "s_per_forward": t.item(), "audio_s_per_s": T / a.sample_rate / t.item()}
if rank == 0 and not a.no_check:
    one = SeparatorEngine(hp, sds, device=dev, mode=a.mode, use_graph=False, small_batch_plan=False)(mix_d)
    out["max_abs_diff_vs_unchunked/rms"] = ((est - one).abs().max() / one.pow(2).mean().sqrt()).item()
if rank == 0:
    print(json.dumps(out), flush=True)
sp.close()          # captured NCCL graphs must go before the process group does
del sp
torch.cuda.synchronize()
if world > 1:
    dist.destroy_process_group()
