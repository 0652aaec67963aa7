"""Same-box GPU comparator for the WHOLE forward: what the reference's module code launches on a GPU, composed from the
libraries this image has -- cuDNN / cuBLAS through torch for the convs and GEMMs, torch elementwise kernels for the
norms, and the sm_100 build of the mamba-ssm selective-scan kernel shipped in the vllm wheel
(``vllm._custom_ops.selective_scan_fwd``) -- against ``SeparatorEngine.forward`` on the same mixtures and weights.

The reference itself cannot run on this box (mamba-ssm / causal-conv1d / speechbrain are not installed);
this script follows its dataflow op for op in the reference's own channel-first layout:

    train_wsj0mix.py:86-111 -> mamba_masknet.py:101-139 -> mamba_blocks.py:186-212 -> bimamba.py:436-462,181-253
    -> selective_scan_interface.py:164-229

with two substitutions, both stated in the output: the depthwise causal conv runs as ``F.conv1d(groups=di)`` + SiLU (the
reference's own non-CUDA branch, bimamba.py:279) because causal_conv1d_cuda is absent, and RMSNorm is plain torch ops
instead of the Triton kernel.  It is a *comparator* (library code, not the product, not the oracle): nothing under
``avse_challenge_b200/`` or ``tests/`` imports it.

    python tools/forward_vs_library.py [--hparams S] [--batch 32] [--seconds 4] [--mode fp32] [--iters 5]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from avse_challenge_b200 import CONFIGS, init_state_dicts, synth_mixture
from avse_challenge_b200.engine import SeparatorEngine

ap = argparse.ArgumentParser()
ap.add_argument("--hparams", default="S"); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--seconds", type=float, default=4.0); ap.add_argument("--sample-rate", type=int, default=8000)
ap.add_argument("--mode", default="fp32"); ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
hp = CONFIGS[a.hparams]
N, D, di, R, Ns = hp.enc_dim, hp.d_model, hp.d_inner, hp.dt_rank, hp.d_state
T = int(a.seconds * a.sample_rate)
dev = "cuda"
io = torch.float32 if a.mode == "fp32" else torch.bfloat16

from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn   # noqa: E402  (library comparator)

sds = init_state_dicts(hp, 1234)
mix_cpu, _ = synth_mixture(a.batch, T, a.sample_rate, seed=1234)
mix = mix_cpu.to(dev)
enc = {k: v.to(dev) for k, v in sds["encoder"].items()}
dec = {k: v.to(dev) for k, v in sds["decoder"].items()}
m = {k: v.to(dev) for k, v in sds["masknet"].items()}
# parameters the scan kernel wants in fp32 (A = -exp(A_log), bimamba.py:227,238); GEMM / conv weights in the I/O dtype
lay = []
for i in range(hp.n_mamba):
    p = f"mamba_net.layers.{i}.mixer."
    dirs = []
    for sfx, a_key, d_key in (("", "A_log", "D"), ("_b", "A_b_log", "D_b")):
        dirs.append(dict(conv_w=m[p + f"conv1d{sfx}.weight"].to(io), conv_b=m[p + f"conv1d{sfx}.bias"].to(io),
                         w_x=m[p + f"x_proj{sfx}.weight"].to(io), w_dt=m[p + f"dt_proj{sfx}.weight"].to(io),
                         dt_bias=m[p + f"dt_proj{sfx}.bias"].float(), A=-torch.exp(m[p + a_key].float()),
                         D=m[p + d_key].float()))
    lay.append(dict(norm=m[f"mamba_net.layers.{i}.norm.weight"], w_in=m[p + "in_proj.weight"].to(io),
                    w_out=m[p + "out_proj.weight"].to(io), dirs=dirs))
state = None


def rmsnorm(x, w, eps=1e-5):                       # mamba_blocks.py:120; fp32 math, cast back
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * w).to(x.dtype)


def inner(xz, q, Bn, L):                           # selective_scan_interface.py:164-229, xz [B, 2di, L]
    x, z = xz.chunk(2, dim=1)
    x = F.silu(F.conv1d(x, q["conv_w"], q["conv_b"], padding=3, groups=di)[..., :L])          # :182 (bimamba.py:279 form)
    x_dbl = F.linear(x.transpose(1, 2).reshape(Bn * L, di), q["w_x"])                          # :186
    delta = (q["w_dt"] @ x_dbl[:, :R].t()).reshape(di, Bn, L).transpose(0, 1).contiguous()    # :187
    Bm = x_dbl[:, R:R + Ns].reshape(Bn, L, Ns).transpose(1, 2).unsqueeze(1).contiguous()      # :198
    Cm = x_dbl[:, R + Ns:].reshape(Bn, L, Ns).transpose(1, 2).unsqueeze(1).contiguous()       # :210
    return selective_scan_fn(x.contiguous(), state, delta, q["A"], Bm, Cm, q["D"], z.contiguous(), q["dt_bias"],
                             delta_softplus=True)                                              # :218-220 (out_z)


@torch.no_grad()
def library_forward(mix):
    global state
    Bn = mix.shape[0]
    mix_w = F.relu(F.conv1d(mix[:, None, :], enc["conv1d.weight"], stride=hp.stride))          # Encoder, [B, N, L]
    L = mix_w.shape[-1]
    if state is None:
        state = torch.zeros(Bn, di, Ns, device=dev)
    x = mix_w.permute(0, 2, 1)                                                                 # mamba_masknet.py:115
    mean = x.mean(-1, keepdim=True); var = x.var(-1, keepdim=True, unbiased=False)             # cLN, :118
    x = (m["layer_norm.gamma"] * (x - mean) / torch.sqrt(var + 1e-8) + m["layer_norm.beta"]).to(io)
    h = F.linear(x, m["bottleneck_conv1x1.conv.weight"][:, :, 0].to(io))                       # :121
    res = None
    for lw in lay:
        res = h if res is None else h + res                                                    # bimamba.py:445-446
        hn = rmsnorm(res, lw["norm"])                                                          # :447
        xz = F.linear(hn, lw["w_in"]).transpose(1, 2)                                          # :192-196, [B, 2di, L]
        out = inner(xz, lw["dirs"][0], Bn, L)                                                  # :221-236
        out_b = inner(xz.flip(-1), lw["dirs"][1], Bn, L)                                       # :237-252
        h = F.linear((0.5 * out + 0.5 * out_b.flip(-1)).transpose(1, 2), lw["w_out"])          # :253
    hn = rmsnorm(h + res, m["mamba_net.norm_f.weight"])                                        # mamba_blocks.py:198-212
    score = F.linear(hn, m["mask_conv1x1.conv.weight"][:, :, 0].to(io))                        # mamba_masknet.py:123
    mask = F.relu(score.float().view(Bn, L, hp.n_spk, N).permute(2, 0, 3, 1))                  # :126-136, [spk, B, N, L]
    sep = torch.stack([mix_w] * hp.n_spk) * mask                                               # train_wsj0mix.py:91-92
    est = torch.cat([F.conv_transpose1d(sep[i], dec["weight"], stride=hp.stride).transpose(1, 2)
                     for i in range(hp.n_spk)], dim=-1)                                        # :95-101, [B, T_est, spk]
    if est.shape[1] < T:                                                                       # :104-109
        est = F.pad(est, (0, 0, 0, T - est.shape[1]))
    return est[:, :T]


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


eng = SeparatorEngine(hp, sds, device=dev, mode=a.mode)
ours = eng.forward(mix)
# parity of the two arms with the library's GEMMs in full fp32 (fp32 mode only; bf16 rounding points differ)
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
lib = library_forward(mix)
rel = ((ours - lib.float()).abs().max() / lib.float().pow(2).mean().sqrt()).item()
res = {"shape": [a.hparams, a.batch, T, a.mode], "audio_s_per_step": a.batch * T / a.sample_rate,
       ("max_abs_diff/rms_vs_library_fp32" if a.mode == "fp32" else "max_abs_diff/rms_vs_library_bf16"): rel}
ms_ours = timed(lambda: eng.forward(mix), a.iters)
res["mtn_forward_ms"] = round(ms_ours, 3)
if a.mode == "fp32":
    res["library_fp32_gemm_ms"] = round(timed(lambda: library_forward(mix), a.iters), 3)
torch.backends.cuda.matmul.allow_tf32 = True      # BASELINE config 2: "tf32 GEMMs"
torch.backends.cudnn.allow_tf32 = True
ms_lib = timed(lambda: library_forward(mix), a.iters)
res["library_%s_gemm_ms" % ("tf32" if a.mode == "fp32" else "bf16")] = round(ms_lib, 3)
res["speedup_vs_library_fastest"] = round(ms_lib / ms_ours, 2)
res["audio_s_per_s"] = {"mtn": round(res["audio_s_per_step"] / ms_ours * 1e3, 1),
                        "library": round(res["audio_s_per_step"] / ms_lib * 1e3, 1)}
res["library"] = ("torch %s (cuBLAS / cuDNN / eager elementwise) + vllm._custom_ops.selective_scan_fwd; causal conv as "
                  "F.conv1d(groups) + SiLU, RMSNorm as torch ops (causal_conv1d_cuda / Triton RMSNorm absent)" % torch.__version__)
print(json.dumps(res), flush=True)
