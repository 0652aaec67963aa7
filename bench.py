#!/usr/bin/env python
"""Benchmark of the Mamba-TasNet separator forward on B200 (BASELINE.json metric: audio-seconds separated per
wall-second; selective-scan HBM GB/s against the measured roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--hparams S] [--batch 32] [--mode fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one forward of the whole separator over one batch of synthetic mixtures.  Default workload =
BASELINE config 2: S hparams, 32 x 4 s @ 8 kHz per GPU, fp32 mode.  N > 1: one process per GPU, each rank separates
its own batch (utterances are independent: batch sharding, no collective on the data path; weak scaling).

One JSON line on rank 0:
  also       (default workload only; --no-also skips it) the two multi-GPU splits BASELINE.json names, a few steps each,
             in the same launch: config 3 (L, bf16 mode, global batch 256 sharded over the ranks: strong scaling) and
             config 5 (one 10-minute 16 kHz recording, sequence-parallel over the ranks) -- each with value / e2e /
             ms_per_step, a parity record measured in this run across the ranks that ran it, and the collectives per forward
  value      whole-job audio-s/s, inputs resident in HBM, CUDA-graph replay, device-timed, max over ranks
  e2e        same metric through the public API with HOST buffers (pinned H2D of the mixtures + D2H of the estimates
             inside the timed region)
  roofline   selective-scan kernel: algorithmic bytes (SURVEY.md 8d / DESIGN.md) / mean launch time measured here
             with CUDA events on the launch stream, vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle port of the reference CPU path (torch loop selective_scan_ref), bounded sample, N=1 only
--impl reference times that CPU path alone (the reference is Python and needs /root/reference + third-party
packages that are absent on the GPU box; the oracle port is the same algorithm, see oracle/restate.py).  Its
config.workload names the BOUNDED SAMPLE it really timed (a CPU forward of the full 32 x 4 s batch takes minutes per
step); config.sample_of names the workload the sample was cut from.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hparams", default="S", choices=["XS", "S", "M", "L"])
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU")
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--sample-rate", type=int, default=8000)
    ap.add_argument("--mode", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--fuse-norm", action="store_true",
                    help="plan with Add -> RMSNorm folded into the GEMM epilogues (one launch fewer per layer)")
    ap.add_argument("--model", default="mambatasnet", choices=["mambatasnet", "dpmamba"],
                    help="dpmamba: the dual-path recipes (hparams/WSJ0Mix/dpmamba_*.yaml), SURVEY 8f rank 1")
    ap.add_argument("--causal", action="store_true",
                    help="unidirectional stack (bidirectional=False, mamba_blocks.py:128): one scan direction per layer")
    ap.add_argument("--chunk-ms", type=float, default=20.0, help="stream: audio per push() call")
    ap.add_argument("--no-fused-push", action="store_true",
                    help="stream: run short chunks through the batch plan (~100 launches) instead of the one-launch cluster kernel")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "longform", "stream", "custom"],
                    help="cfg2 (default): S, 32 x 4 s @ 8 kHz per GPU, fp32 mode, weak scaling.  cfg3: L, global batch "
                         "256 x 4 s sharded over the GPUs, bf16 mode, strong scaling.  longform: one 10-minute 16 kHz "
                         "recording, sequence-parallel over the GPUs (S fp32 unless --hparams/--mode given), strong "
                         "scaling.  custom: take --hparams/--batch/--seconds/--sample-rate/--mode as given.")
    ap.add_argument("--sub-chunks", type=int, default=74,
                    help="longform: time sub-chunks per GPU (74 x 32 warp pairs = two full waves of a 148-SM GPU for S hparams)")
    ap.add_argument("--exchange", default="allgather", choices=["allgather", "sendrecv"])
    ap.add_argument("--no-also", action="store_true", help="default workload: skip the config-3 / config-5 side runs")
    ap.add_argument("--also-steps", type=int, default=3, help="timed steps of each side run")
    a = ap.parse_args()
    explicit = {x.split("=")[0] for x in sys.argv[1:] if x.startswith("--")}
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.workload == "cfg3":
        if "--hparams" not in explicit: a.hparams = "L"
        if "--mode" not in explicit: a.mode = "bf16"
        if "--batch" not in explicit: a.batch = max(1, 256 // world)
    elif a.workload == "cfg4":   # AVSEC-4-shaped: 16 kHz mono 6 s noisy mixtures, 4096 utterances over the GPUs in micro-batches
        if "--seconds" not in explicit: a.seconds = 6.0
        if "--sample-rate" not in explicit: a.sample_rate = 16000
        if "--mode" not in explicit: a.mode = "bf16"
        if "--batch" not in explicit: a.batch = 64
    elif a.workload == "longform":
        if "--seconds" not in explicit: a.seconds = 600.0
        if "--sample-rate" not in explicit: a.sample_rate = 16000
        a.batch = 1
    elif a.workload == "stream":
        a.causal = True
        if "--batch" not in explicit: a.batch = 1
    return a


def workload_config(a, n_gpus):
    prec = ("(split-bf16 x3 tcgen05 GEMMs, fp32 scan state)" if a.mode == "fp32"
            else "(bf16 tcgen05 GEMMs + bf16 activations, fp32 scan state)")
    name = {"cfg2": "BASELINE config 2", "cfg3": "BASELINE config 3", "longform": "BASELINE config 5",
            "cfg4": "BASELINE config 4 (one micro-batch per step; 4096 utterances = 4096 / (batch x GPUs) steps)",
            "custom": "custom", "stream": "streaming (SURVEY 8f rank 2)"}[a.workload]
    if a.causal:
        name += " [causal: bidirectional=False]"
    if a.model == "dpmamba":
        name += " [DPMamba mask network: Dual_Path_Model, K=250]"
    if a.workload == "longform":
        return {
            "workload": (f"{name}: Mamba-TasNet {a.hparams} hparams, one {a.seconds:g} s @ {a.sample_rate // 1000} kHz "
                         f"recording, {a.mode} mode {prec}, time sharded over {n_gpus} GPU(s) x {a.sub_chunks} sub-chunks"),
            "hparams": a.hparams, "batch_per_gpu": 1, "global_batch": 1, "seconds": a.seconds,
            "sample_rate": a.sample_rate, "mode": a.mode,
            "parallelism": f"sequence-parallel x{n_gpus}: conv halo + chunk-summary exchange ({a.exchange}) per layer",
            "l2": "no flush needed: every layer streams far more than the 126 MB L2",
        }
    return {
        "workload": (f"{name}: Mamba-TasNet {a.hparams} hparams, {a.batch} x {a.seconds:g} s @ "
                     f"{a.sample_rate // 1000} kHz 2-speaker mixtures per GPU, {a.mode} mode {prec}"),
        "hparams": a.hparams, "batch_per_gpu": a.batch, "global_batch": a.batch * n_gpus, "seconds": a.seconds,
        "sample_rate": a.sample_rate, "mode": a.mode,
        "parallelism": f"batch-sharded x{n_gpus}, no collective on the data path",
        "l2": "no flush needed: every step streams >3 GB of activations per layer group, far larger than the 126 MB L2",
    }


def scaling_kind(a):
    return "weak" if a.workload in ("cfg2", "cfg4", "custom", "stream") else "strong"


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_throughput(hparams: str, sample_rate: int, steps: int, warmup: int, budget_s: float, causal: bool = False,
                             model: str = "mambatasnet"):
    """Time the oracle port of the reference's CPU forward (selective_scan_ref = per-step torch loop,
    Mamba-TasNet/modules/mamba/selective_scan_interface.py:91-157) on a bounded sample of the workload."""
    import torch
    from avse_challenge_b200 import CONFIGS, init_state_dicts, synth_mixture
    from oracle import restate

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if model == "dpmamba":
        from avse_challenge_b200 import DP_CONFIGS, init_dp_state_dicts
        hp = DP_CONFIGS[hparams]
        sds = init_dp_state_dicts(hp, 1234)
    else:
        hp = CONFIGS[hparams].causal() if causal else CONFIGS[hparams]
        sds = init_state_dicts(hp, 1234)

    def run(seconds):
        T = int(round(seconds * sample_rate)) // 8 * 8
        mix, _ = synth_mixture(1, T, sample_rate, seed=1234)
        t0 = time.perf_counter()
        with torch.no_grad():
            if model == "dpmamba":
                restate.separate_dp(mix, sds, hp, scan_impl="torch")
            else:
                restate.separate(mix, sds, hp.n_mamba, scan_impl="torch")
        return time.perf_counter() - t0, T / sample_rate

    # calibrate on a very short clip, then pick the longest sample that keeps (steps + warmup) inside the budget
    t_cal, s_cal = run(0.125)
    per_audio_s = t_cal / s_cal
    n = max(1, steps + warmup)
    seconds = 4.0
    while seconds > 0.125 and per_audio_s * seconds * n > budget_s:
        seconds /= 2
    for _ in range(warmup):
        run(seconds)
    times = []
    for _ in range(max(1, steps)):
        t, s = run(seconds)
        times.append(t)
    dt = sum(times) / len(times)
    return {"value": s / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample_seconds": seconds,
            "sample": f"1 utterance x {seconds:g} s @ {sample_rate} Hz, {hparams} hparams, fp32, torch {torch.get_num_threads()} threads, "
                      f"{len(times)} timed run(s) of {dt:.2f} s"}, dt


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, dt = cpu_reference_throughput(a.hparams, a.sample_rate, a.steps, a.warmup, budget_s=150.0, causal=a.causal, model=a.model)
    cfg = workload_config(a, a.gpus)
    # say what was really timed: a bounded sample of the workload (one short utterance), not the batch the GPU arm runs
    cfg["sample_of"] = cfg["workload"]
    cfg["workload"] = f"BOUNDED SAMPLE ({base['sample']}) of: {cfg['sample_of']}"
    cfg["batch_per_gpu"], cfg["global_batch"], cfg["seconds"] = 1, 1, base["sample_seconds"]
    cfg["parallelism"] = f"CPU, torch with {base['cores']} threads, no GPU"
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": scaling_kind(a),
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.f.read().splitlines():
            c = [x.strip() for x in row.split(",")]
            if len(c) < 8:
                continue
            try:
                clk, mx, util = float(c[0]), float(c[1]), float(c[3])
            except ValueError:
                continue
            smax.append(mx)
            if util >= 50:
                sm.append(clk)
                for n, v in zip(names, c[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples_under_load": len(sm)}


# ------------------------------------------------------------------------------------------ B200 arm
def scan_algorithmic_bytes(hp, batch, L, mode, ndir=2):
    """Bytes one both-direction scan launch must move at the op boundary of the reference's
    selective_scan_cuda.fwd: per direction B*L*[(u, delta, z read + out written)*di + (B, C)*Ns]*s_io
    + parameters.  fp32 mode moves 4-byte elements (u / y as hi+lo bf16 planes = 4 B), bf16 mode 2-byte."""
    s_io = 4 if mode == "fp32" else 2
    di, Ns = hp.d_inner, hp.d_state
    per_dir = batch * L * (4 * di + 2 * Ns) * s_io + (di * Ns + 2 * di) * 4
    return ndir * per_dir


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", d
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def scan_source_hash() -> str:
    """Hash of the sources the scan kernel is compiled from: a committed DRAM-traffic figure (tools/scan_traffic.py) is
    only quoted while it belongs to this code."""
    import hashlib
    h = hashlib.sha256()
    for f in ("mtn_scan.cu", "mtn_scan_pair.cuh", "mtn_ptx.cuh"):
        with open(os.path.join(ROOT, "avse_challenge_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def measured_traffic(key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one scan launch of this workload, from the ncu capture that
    tools/scan_traffic.py wrote to profiles/scan_traffic.json -- or None when no capture of THIS kernel source exists."""
    tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
    try:
        tj = json.load(open(tp))
    except Exception:
        return None, "no profiles/scan_traffic.json"
    if tj.get("source_hash") != scan_source_hash():
        return None, "stale: profiles/scan_traffic.json was captured for another revision of the scan sources"
    ent = tj.get("launches", {}).get(key)
    if ent is None:
        return None, f"no capture of {key}"
    return int(ent["dram_bytes_per_launch"]), ent.get("source", "tools/scan_traffic.py")


class Ctx:
    """Process-group plumbing shared by the timed arms of one bench.py process."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product path)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def gather(self, t):
        """[*shape] on every rank -> [world, *shape] on every rank."""
        if self.world == 1:
            return t.unsqueeze(0)
        out = self.torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1))
        return out

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed(ctx, fn, steps, warmup, finish=None):
    """`warmup` untimed calls, then exactly `steps` calls between CUDA events on the launch stream, barrier + synchronize on
    both sides; returns the max over ranks of the elapsed ms.  `finish` (optional) runs after the last call, inside the timed
    region: it makes the launch stream wait for work the calls left on other streams (pipelined host copies)."""
    torch = ctx.torch
    for _ in range(warmup):
        fn()
    if finish is not None:
        finish()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    if finish is not None:
        finish()
    e1.record()
    ctx.barrier()
    return e0.elapsed_time(e1)


def measure_batch(a, ctx, steps, warmup, cpu_baseline=True, parity=False, sampler=None):
    """One batch-sharded workload (configs 2 / 3 / 4 / custom): every rank separates its own `a.batch` utterances."""
    torch = ctx.torch
    from avse_challenge_b200 import CONFIGS, init_state_dicts, synth_mixture
    from avse_challenge_b200.engine import SeparatorEngine
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    T = int(round(a.seconds * a.sample_rate)) // 8 * 8
    if a.model == "dpmamba":
        from avse_challenge_b200 import DP_CONFIGS, init_dp_state_dicts
        from avse_challenge_b200.dpmamba import DPSeparatorEngine
        from avse_challenge_b200 import ops as _ops
        dhp = DP_CONFIGS[a.hparams]
        hp = dhp.stack                        # what one scan launch sees: the intra / inter stack
        sds = init_dp_state_dicts(dhp, 1234)
        L = dhp.frames(T)
        scan_tokens = a.batch * _ops.dp_num_chunks(L, dhp.chunk_size) * dhp.chunk_size
    else:
        hp = CONFIGS[a.hparams].causal() if a.causal else CONFIGS[a.hparams]
        L = hp.frames(T)
        sds = init_state_dicts(hp, 1234)
        scan_tokens = a.batch * L
    mix_cpu, _ = synth_mixture(min(a.batch, 8), T, a.sample_rate, seed=1234 + rank,
                               noise_second_source=(a.workload == "cfg4"))    # cfg4: speech + noise at 0 dB
    reps = (a.batch + mix_cpu.shape[0] - 1) // mix_cpu.shape[0]
    mix_cpu = mix_cpu.repeat(reps, 1)[: a.batch].contiguous()      # synthetic batch (8 distinct voices tiled)
    if a.model == "dpmamba":
        eng = DPSeparatorEngine(dhp, sds, device=dev, mode=a.mode, use_graph=not a.no_graph)
    else:
        eng = SeparatorEngine(hp, sds, device=dev, mode=a.mode, use_graph=not a.no_graph, fuse_norm=a.fuse_norm)
    audio_s_per_step = a.batch * T / a.sample_rate

    cpu_base = None
    if cpu_baseline and rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu_base, _ = cpu_reference_throughput(a.hparams, a.sample_rate, steps=1, warmup=0, budget_s=30.0, causal=a.causal, model=a.model)

    # ---- parity record of this run: the SAME two utterances on every rank -> identical results across the GPUs, and
    # inside the full batch == run alone (utterances are independent, so any difference is a kernel / sharding bug)
    par = None
    if parity:
        pmix, _ = synth_mixture(2, T, a.sample_rate, seed=4321)
        sbp = getattr(eng, "small_batch_plan", False)
        if sbp:
            eng.small_batch_plan = False          # "alone" = the batch plan at B = 2, the kernels the full batch runs
        alone = eng.forward(pmix.to(dev))
        d_chunked = None
        if sbp:
            eng.small_batch_plan = True           # what forward() picks for so small a batch: the chunked-scan plan
            chunked = eng.forward(pmix.to(dev))
            d_chunked = ((chunked - alone).abs().max() / alone.pow(2).mean().sqrt().clamp(min=1e-30)).item()
        ws0 = eng.workspace(a.batch, T)
        ws0.mix[:, :T].copy_(mix_cpu.to(dev))
        ws0.mix[:2, :T].copy_(pmix.to(dev))
        inside = eng.forward_into_workspace(a.batch, T)[:2].clone()
        rms = alone.pow(2).mean().sqrt().clamp(min=1e-30)
        d_batch = ((inside - alone).abs().max() / rms).item()
        allr = ctx.gather(alone)
        d_ranks = ((allr - allr[0:1]).abs().max() / rms).item()
        par = {"max_abs_over_rms": max(d_batch, d_ranks), "across_ranks": d_ranks, "batch_vs_alone": d_batch, "world": world,
               "check": "the same 2 utterances separated by every rank's engine, compared across the ranks and, on each "
                        "rank, inside the full batch against run alone (batch plan both times)"}
        if d_chunked is not None:
            par["small_batch_chunked_plan_vs_batch_plan"] = d_chunked

    # ---- device-resident timing (value): inputs already in HBM, graph replay
    ws = eng.workspace(a.batch, T)
    ws.mix[:, :T].copy_(mix_cpu.to(dev))
    ms_total = timed(ctx, lambda: eng.forward_into_workspace(a.batch, T), steps, max(3, warmup))

    # ---- end-to-end through the public API with host buffers
    pin_in = mix_cpu.pin_memory()
    pin_out = torch.empty((a.batch, T, hp.n_spk), dtype=torch.float32).pin_memory()
    dmix = torch.empty((a.batch, T), dtype=torch.float32, device=dev)

    if hasattr(eng, "forward_host"):
        # the serving call: pinned host mixtures in, pinned host estimates out, copies of neighbouring calls overlapped with the
        # kernels (two staging buffers, two copy streams); the timed region ends when the LAST estimate has reached host memory
        pin_outs = [pin_out, torch.empty_like(pin_out).pin_memory()]
        step_no = [0]

        def e2e_step():
            eng.forward_host(pin_in, pin_outs[step_no[0] & 1])
            step_no[0] += 1

        ms_e2e = timed(ctx, e2e_step, steps, 3, finish=eng.wait_host)
    else:
        def e2e_step():
            dmix.copy_(pin_in, non_blocking=True)
            est = eng.forward(dmix)
            pin_out.copy_(est, non_blocking=True)

        ms_e2e = timed(ctx, e2e_step, steps, 3)

    # ---- per-kernel timing in situ (CUDA events around every launch, eager mode), for the roofline
    prof = eng.profile_ops(a.batch, T, steps=max(1, min(steps, 5)))
    clocks = sampler.stop() if sampler else None
    ms_total, ms_e2e = ctx.max_over_ranks(ms_total, ms_e2e)
    del eng, ws, dmix
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src, peaks = load_peaks()
    scan_ms = prof["scan"]["ms"]
    alg = scan_algorithmic_bytes(hp, 1, scan_tokens, a.mode, 2 if hp.bidirectional else 1)
    achieved = alg / (scan_ms * 1e-3) / 1e9
    tkey = f"{a.hparams}_b{a.batch}_L{L}_{a.mode}" + ("" if hp.bidirectional else "_causal") + ("_dp" if a.model == "dpmamba" else "")
    traffic, traffic_src = measured_traffic(tkey)
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    n_exp = (2 if hp.bidirectional else 1) * scan_tokens * hp.d_inner * hp.d_state
    mufu_ms = n_exp / (148 * 16 * sm_mhz * 1e6) * 1e3
    step_ms = ms_total / steps
    total_prof = sum(v["ms_per_forward"] for v in prof.values())
    line = {
        "metric": METRIC, "value": audio_s_per_step * world / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": max(3, warmup), "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": scaling_kind(a), "vs_baseline": None, "dtype": "f32" if a.mode == "fp32" else "bf16",
        "data": "synthetic", "config": workload_config(a, world),
        "e2e": {"value": audio_s_per_step * world / (ms_e2e / steps * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": a.batch * T * 4, "d2h_bytes_per_step": a.batch * T * hp.n_spk * 4,
                "ms_per_step": ms_e2e / steps},
        "gpu_launches": steps * sum(v["launches"] for v in prof.values()) + steps,  # +1: decoder = 2 kernels
        "launches_per_step": sum(v["launches"] for v in prof.values()) + 1,
        "roofline": {"kernel": "mtn::scan_kernel_pair (selective scan, both directions per launch)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": alg,
                     "ms_per_launch": scan_ms, "launches_per_step": prof["scan"]["launches"],
                     "share_of_step": prof["scan"]["ms_per_forward"] / total_prof,
                     "mufu_bound_ms": mufu_ms, "mufu_frac": mufu_ms / scan_ms},
        "kernels_ms_per_step": {k: round(v["ms_per_forward"], 4) for k, v in sorted(prof.items())},
        "clocks": clocks,
    }
    if par is not None:
        line["parity"] = par
        line["collectives_per_forward"] = 0
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    return line


def measure_longform(a, ctx, steps, warmup, cpu_baseline=True, parity=False, sampler=None):
    """BASELINE config 5: one long recording, sequence-parallel over the ranks (strong scaling)."""
    torch = ctx.torch
    from avse_challenge_b200 import CONFIGS, init_state_dicts, synth_mixture
    from avse_challenge_b200.parallel import SequenceParallelSeparator
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    hp = CONFIGS[a.hparams]
    T = int(round(a.seconds * a.sample_rate)) // 8 * 8
    sds = init_state_dicts(hp, 1234)
    mix_cpu, _ = synth_mixture(1, min(T, 30 * a.sample_rate), a.sample_rate, seed=1234)
    reps = -(-T // mix_cpu.shape[1])
    mix_cpu = mix_cpu.repeat(1, reps)[:, :T].contiguous()            # 30 s of synthetic speech tiled to length
    sp = SequenceParallelSeparator(hp, sds, device=dev, mode=a.mode, sub_chunks=a.sub_chunks, exchange=a.exchange,
                                   use_graph=not a.no_graph)
    cpu_base = None
    if cpu_baseline and rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu_base, _ = cpu_reference_throughput(a.hparams, a.sample_rate, steps=1, warmup=0, budget_s=30.0, causal=a.causal, model=a.model)

    # ---- parity record of this run: a short clip through the sequence-parallel path on all ranks (collectives and all)
    # against the single-GPU batch plan (SeparatorEngine, itself oracle-tested) on every rank
    par = None
    if parity:
        from avse_challenge_b200.engine import SeparatorEngine
        Tp = min(T, 12 * a.sample_rate) // 8 * 8
        clip = mix_cpu[:, :Tp].contiguous().to(dev)
        got = sp(clip)
        ref = SeparatorEngine(hp, sds, device=dev, mode=a.mode, use_graph=False, small_batch_plan=False)(clip)
        rms = ref.pow(2).mean().sqrt().clamp(min=1e-30)
        err = ((got - ref).abs().max() / rms).item()
        err, = ctx.max_over_ranks(err)
        par = {"max_abs_over_rms": err, "world": world, "clip_seconds": Tp / a.sample_rate,
               "check": "sequence-parallel forward of a short clip over all ranks (replicated output) against the one-GPU "
                        "batch plan (SeparatorEngine) on each rank, max over ranks"}
        del ref, got
        torch.cuda.empty_cache()

    # ---- sharded timing: every rank holds only its slice of the recording and produces only its slice of the estimate
    s0, s1 = sp.input_range(T)
    o0, o1 = sp.output_range(T)
    mix_loc = mix_cpu[:, s0:s1].contiguous()
    mix_d = mix_loc.to(dev)
    ms_total = timed(ctx, lambda: sp.forward_local(mix_d, T), steps, max(3, warmup))
    pin_in = mix_loc.pin_memory()
    pin_out = torch.empty((o1 - o0, hp.n_spk), dtype=torch.float32).pin_memory()
    dmix = torch.empty((1, s1 - s0), dtype=torch.float32, device=dev)

    def e2e_step():
        dmix.copy_(pin_in, non_blocking=True)
        pin_out.copy_(sp.forward_local(dmix, T), non_blocking=True)

    ms_e2e = timed(ctx, e2e_step, steps, 2)
    # ---- the two scan passes timed in situ: one eager forward on every rank (the collectives need all of them) with CUDA
    # events around the summary pass and the seeded pass of every layer
    be, spans = sp.be, {"summary": [], "seeded": []}

    def _spanned(fn, key):
        def call(*args, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*args, **kw)
            e1.record()
            spans[key].append((e0, e1))
            return out
        return call

    orig = (be.scan_summary, be.scan_seeded)
    be.scan_summary, be.scan_seeded = _spanned(orig[0], "summary"), _spanned(orig[1], "seeded")
    try:
        for it in range(2):                                               # first pass warms the eager path up
            for v in spans.values():
                v.clear()
            sp._forward_local(mix_d, T)
            torch.cuda.synchronize()
    finally:
        be.scan_summary, be.scan_seeded = orig
    scan_ms = {k: sum(e0.elapsed_time(e1) for e0, e1 in v) for k, v in spans.items()}
    n_scan_layers = len(spans["seeded"])
    scan_sum_ms, scan_seed_ms = ctx.max_over_ranks(scan_ms["summary"], scan_ms["seeded"])
    clocks = sampler.stop() if sampler else None
    ms_total, ms_e2e = ctx.max_over_ranks(ms_total, ms_e2e)
    graphed = bool(sp.use_graph and sp.graph_failed is None)
    graph_note = sp.graph_failed
    ncoll = sp.collectives_per_forward
    del sp, dmix, mix_d
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    from avse_challenge_b200.parallel import make_seq_plan
    peak, peak_src, _ = load_peaks()
    audio_s = T / a.sample_rate
    L = hp.frames(T)
    plan = make_seq_plan(L, world, a.sub_chunks)
    step_ms = ms_total / steps
    alg = scan_algorithmic_bytes(hp, 1, plan.ranges[0][1] - plan.ranges[0][0], a.mode, 2 if hp.bidirectional else 1)
    per_layer_ms = (scan_sum_ms + scan_seed_ms) / max(1, n_scan_layers)
    achieved = alg / (per_layer_ms * 1e-3) / 1e9
    per_fwd = 4 + hp.n_mamba * 10     # enc, bottleneck, (norm, in_proj, 2 edge copies, conv, x_proj, scan A, fold, scan B, out_proj) x layers, norm_f+mask, decoder(2)
    line = {
        "metric": METRIC, "value": audio_s / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": max(3, warmup), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32" if a.mode == "fp32" else "bf16", "data": "synthetic",
        "config": dict(workload_config(a, world), frames=L, frames_per_gpu=plan.ranges[0][1] - plan.ranges[0][0],
                       sub_chunk_frames=plan.Ls),
        "e2e": {"value": audio_s / (ms_e2e / steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": (s1 - s0) * 4,
                "d2h_bytes_per_step": (o1 - o0) * hp.n_spk * 4, "ms_per_step": ms_e2e / steps,
                "note": "per rank: every rank copies only its own slice of the recording in and of the estimate out"},
        "gpu_launches": steps * (per_fwd + 1),
        "collectives_per_forward": ncoll, "cuda_graph": graphed,
        "roofline": {"kernel": "mtn::scan_kernel (summary pass) + mtn::scan_kernel_pair (seeded pass) per layer", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "traffic": None, "algorithmic_bytes_per_launch": alg, "ms_per_launch": per_layer_ms,
                     "ms_summary_pass": scan_sum_ms / max(1, n_scan_layers), "ms_seeded_pass": scan_seed_ms / max(1, n_scan_layers),
                     "launches_per_step": 2 * n_scan_layers, "share_of_step": (scan_sum_ms + scan_seed_ms) / step_ms,
                     "note": "algorithmic bytes = op-boundary bytes of ONE scan over this rank's frames (the summary pass is "
                             "the plan's overhead, not algorithmic traffic) / (summary + seeded pass time of one layer), timed "
                             "with CUDA events in one eager forward, max over ranks; share_of_step relates those eager "
                             "durations to the graph-replayed step"},
        "clocks": clocks,
    }
    if graph_note:
        line["cuda_graph_note"] = graph_note
    if par is not None:
        line["parity"] = par
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    return line


def side_args(a, workload):
    """argparse namespace of a side run (`also`): the named BASELINE config with its defaults."""
    import copy
    b = copy.copy(a)
    b.workload = workload
    world = int(os.environ.get("WORLD_SIZE", "1"))
    b.model, b.causal, b.fuse_norm = "mambatasnet", False, False
    if workload == "cfg3":
        b.hparams, b.mode, b.batch, b.seconds, b.sample_rate = "L", "bf16", max(1, 256 // world), 4.0, 8000
    else:
        b.hparams, b.mode, b.batch, b.seconds, b.sample_rate = "S", "fp32", 1, 600.0, 16000
    return b


def run_b200_arm(a):
    ctx = Ctx()
    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    line = measure_batch(a, ctx, a.steps, a.warmup, sampler=sampler, parity=(ctx.world > 1))
    if a.workload == "cfg2" and a.model == "mambatasnet" and not a.causal and not a.no_also:
        also = []
        for wl, fn in (("cfg3", measure_batch), ("longform", measure_longform)):
            try:
                sub = fn(side_args(a, wl), ctx, max(1, a.also_steps), 1, cpu_baseline=False, parity=True)
            except Exception as e:   # a side run must never take the headline line down with it
                sub = {"config": {"workload": wl}, "error": f"{type(e).__name__}: {e}"}
            also.append(sub)
        if ctx.rank == 0:
            line["also"] = also
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()


def run_longform_arm(a):
    ctx = Ctx()
    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    line = measure_longform(a, ctx, a.steps, a.warmup, sampler=sampler, parity=True)
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()


def run_stream_arm(a):
    """Streaming causal separation: `batch` concurrent streams per GPU, one push() per `chunk_ms` of
    audio.  value = audio-s/s with the chunks already in HBM; e2e = every chunk copied from pinned host memory and its
    estimate copied back (what a live pipeline does); latency = device time of one push."""
    import torch
    import torch.distributed as dist
    from avse_challenge_b200 import CONFIGS, init_state_dicts, synth_mixture
    from avse_challenge_b200.engine import SeparatorEngine
    from avse_challenge_b200.streaming import StreamingSeparator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    hp = CONFIGS[a.hparams].causal()
    n = max(16, int(round(a.chunk_ms * 1e-3 * a.sample_rate)) // 8 * 8)     # samples per push
    sds = init_state_dicts(hp, 1234)
    eng = SeparatorEngine(hp, sds, device=dev, mode=a.mode, use_graph=False, fuse_norm=a.fuse_norm)
    st = StreamingSeparator(eng, a.batch, use_graph=not a.no_graph, fused=False if a.no_fused_push else None)
    fused = st._fused is not None and n // 8 <= 32
    n_chunks = 64
    mix_cpu, _ = synth_mixture(min(a.batch, 8), n * n_chunks, a.sample_rate, seed=1234 + rank)
    mix_cpu = mix_cpu.repeat(-(-a.batch // mix_cpu.shape[0]), 1)[: a.batch].contiguous()
    chunks_d = [mix_cpu[:, i * n:(i + 1) * n].contiguous().to(dev) for i in range(n_chunks)]
    chunks_h = [mix_cpu[:, i * n:(i + 1) * n].contiguous().pin_memory() for i in range(n_chunks)]
    out_h = torch.empty((a.batch, n, hp.n_spk), dtype=torch.float32).pin_memory()
    stage = torch.empty((a.batch, n), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    st.push(chunks_d[0])                                            # first chunk has its own shape (no carried samples)
    for i in range(max(3, a.warmup)):
        st.push(chunks_d[1 + i % (n_chunks - 1)])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        st.push(chunks_d[1 + i % (n_chunks - 1)])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)

    def e2e_step(i):
        stage.copy_(chunks_h[1 + i % (n_chunks - 1)], non_blocking=True)
        out_h.copy_(st.push(stage), non_blocking=True)

    for i in range(3):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(a.steps):
        e2e_step(i)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    # host-observed latency of one live chunk (copy in, push, copy out, wait)
    lat = []
    for i in range(min(a.steps, 50)):
        t0 = time.perf_counter()
        e2e_step(i)
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()
    if rank == 0:
        peak, peak_src, _ = load_peaks()
        audio_s = a.batch * n / a.sample_rate
        step_ms = ms_total / a.steps
        cfg = workload_config(a, world)
        cfg["workload"] = (f"streaming causal Mamba-TasNet {a.hparams} hparams (bidirectional=False), {a.batch} concurrent stream(s) "
                           f"per GPU, {n / a.sample_rate * 1e3:g} ms ({n // 8} frames) per push, {a.mode} mode, "
                           + ("one cluster-kernel launch per push (mtn_stream_push_fwd)" if fused else "CUDA graph per chunk shape"))
        cfg["chunk_samples"] = n
        cfg["plan"] = "fuse_norm (Add -> RMSNorm folded into the GEMM epilogues)" if a.fuse_norm else "separate add_rmsnorm kernel"
        line = {
            "metric": METRIC, "value": audio_s * world / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(3, a.warmup), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if a.mode == "fp32" else "bf16", "data": "synthetic", "config": cfg,
            "e2e": {"value": audio_s * world / (ms_e2e / a.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": a.batch * n * 4,
                    "d2h_bytes_per_step": a.batch * n * hp.n_spk * 4, "ms_per_step": ms_e2e / a.steps},
            "latency_ms": {"chunk_audio_ms": n / a.sample_rate * 1e3, "device_ms_per_push": step_ms,
                           "host_observed_median_ms": statistics.median(lat), "host_observed_max_ms": max(lat),
                           "algorithmic_latency_ms": 16 / a.sample_rate * 1e3},
            "gpu_launches": a.steps * (1 if fused else eng.launches_per_forward),
            "roofline": {"kernel": ("mtn::stream_push_kernel: one cluster per stream, bound by streaming the weights (fp32-equivalent bytes "
                                    "of every GEMM weight once per cluster) from L2 and by its cluster barriers, not by HBM") if fused
                         else "launch-latency bound (about 100 launches of a few-frame chunk per push)", "bound": "hbm",
                         "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "peak_source": peak_src, "traffic": None,
                         "note": "per-kernel roofline is reported on the cfg2 workload"},
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    elif a.workload == "stream":
        run_stream_arm(a)
    elif a.workload == "longform":
        run_longform_arm(a)
    else:
        run_b200_arm(a)


if __name__ == "__main__":
    main()
