"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of the reference Mamba-TasNet forward.

Channel-last (``[B, L, C]``) restatement of the math in SURVEY.md App. A.  It is the checker
for the CUDA path and the reported CPU baseline; the product never imports it.
Pinned against the real reference (imported under ``oracle/ref_shims.py``) by
``tests/test_oracle.py`` via ``tests/golden/*.npz``.

All parameters are taken from state_dicts with the reference's key names.
Citations are relative to ``/root/reference``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB = None


# ----------------------------------------------------------------------------- C scan (optional)
def build_c_oracle(force: bool = False) -> str:
    """Compile ``oracle/selscan_ref.c`` -> ``oracle/_build/libselscan_ref.so`` (gcc, OpenMP)."""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libselscan_ref.so")
    src = os.path.join(_HERE, "selscan_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-fno-fast-math",
                               "-o", so, src, "-lm"])
    return so


def _clib():
    global _CLIB
    if _CLIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        lib.selscan_ref_f32.restype = None
        lib.selscan_ref_f32.argtypes = [ctypes.c_void_p] * 10 + [ctypes.c_int] * 5 + [ctypes.c_void_p] * 2
        _CLIB = lib
    return _CLIB


# ----------------------------------------------------------------------------- precision models
# "fp32"          the reference's fp32 path (selective_scan_ref in fp32): the oracle of BASELINE configs 1-2.
# "product_bf16"  rounds to bf16 exactly where the CUDA path's bf16 mode stores bf16: both operands of
#                 every dense contraction (xn / yn / u / y planes, weights), the x half and the silu(z) half of in_proj's
#                 output, the conv output u, the gated scan output y.  Residual stream, [dt|B|C], delta, the SSM state,
#                 mix_w, the mask product and the decoder stay fp32.  This is the checker of configs 3-4: the product
#                 must agree with it far more tightly than with the fp32 oracle.
# "autocast_ref"  rounds where the REFERENCE's mixed-precision path rounds (``precision: bf16`` is the recipes' default,
#                 hparams/WSJ0Mix/mambatasnet_S.yaml:38; autocast region train_wsj0mix.py:161-164): every conv / linear /
#                 matmul takes bf16 operands and returns bf16 (incl. x_proj and dt_proj, whose weights
#                 selective_scan_interface.py:174-176 casts), the depthwise conv and the scan kernels take bf16 and return
#                 bf16 with fp32 arithmetic inside (A, D, delta_bias and the state stay fp32: bimamba.py:200,232-233),
#                 RMSNorm returns bf16, and the residual stream is bf16 (residual_in_fp32: False).  Pinned against a run
#                 of the reference's own modules under torch.autocast (tests/golden/forward_tiny_autocast.npz).
_PREC = "fp32"
_GEMM_MODE = "fp32"


def set_precision(mode: str):
    """Select one of the three precision models above (also sets the GEMM operand rounding)."""
    global _PREC
    assert mode in ("fp32", "product_bf16", "autocast_ref")
    _PREC = mode
    set_gemm_mode("fp32" if mode == "fp32" else "bf16")


def _r(x):
    """Round to bf16 (round-to-nearest-even), keep the container dtype."""
    return x.to(torch.bfloat16).to(x.dtype)


def _r_ref(x):
    return _r(x) if _PREC == "autocast_ref" else x


def _r_prod(x):
    return _r(x) if _PREC == "product_bf16" else x


def _r_any(x):
    return _r(x) if _PREC != "fp32" else x


def set_gemm_mode(mode: str):
    """Rounding model applied to BOTH operands of every dense contraction of the path:
    "fp32" (reference), "bf16" (one bf16 plane), "bf16x3" (hi+lo bf16 planes, 3 products --
    what the CUDA fp32 mode computes), "tf32" (round-to-nearest 10-bit mantissa)."""
    global _GEMM_MODE
    assert mode in ("fp32", "bf16", "bf16x3", "tf32")
    _GEMM_MODE = mode


def _planes(x):
    hi = x.to(torch.bfloat16).to(x.dtype)
    lo = (x - hi).to(torch.bfloat16).to(x.dtype)
    return hi, lo


def _mm(x, w):
    """``x @ w.T`` under the selected operand-rounding model (accumulation stays fp32/fp64)."""
    if _GEMM_MODE == "fp32":
        return x @ w.t()
    if _GEMM_MODE == "bf16":
        return _r_ref(_r(x) @ _r(w).t())          # autocast: the output is bf16 too; the CUDA path keeps fp32 outputs
    if _GEMM_MODE == "tf32":
        r = lambda t: ((t.float().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32).to(t.dtype)
        return r(x) @ r(w).t()
    xh, xl = _planes(x)
    wh, wl = _planes(w)
    return xh @ wh.t() + (xl @ wh.t() + xh @ wl.t())


# ----------------------------------------------------------------------------- pieces
def encoder_fwd(mix, w_enc):
    """``relu(conv1d(mix[:,None,:], W[N,1,K], stride K//2))`` -> ``[B, L, N]``.
    speechbrain ``dual_path.Encoder`` == ``baseline/avse2/model.py:14-24``."""
    k = w_enc.shape[-1]
    if _PREC == "autocast_ref":   # conv1d is an autocast op: bf16 operands, bf16 result
        return F.relu(_r(F.conv1d(_r(mix).unsqueeze(1), _r(w_enc), stride=k // 2))).transpose(1, 2).contiguous()
    return F.relu(F.conv1d(mix.unsqueeze(1), w_enc, stride=k // 2)).transpose(1, 2).contiguous()


def cln_fwd(y, gamma, beta, eps=1e-8):
    """ChannelwiseLayerNorm over the channel axis, biased variance, eps 1e-8
    (speechbrain ``conv_tasnet.ChannelwiseLayerNorm``; ctor ``modules/mamba_masknet.py:73``)."""
    if _PREC == "autocast_ref":
        # y is bf16: mean / var / (y - mean) / var + eps / pow are bf16 ops (fp32 inside, one rounding each); the products
        # with the fp32 parameters gamma / beta promote to fp32 (type promotion with a dimensioned fp32 tensor)
        mean = _r(y.mean(dim=-1, keepdim=True))
        var = _r(y.var(dim=-1, keepdim=True, unbiased=False))
        return gamma.reshape(-1) * _r(y - mean) / _r(torch.sqrt(_r(var + eps))) + beta.reshape(-1)
    mean = y.mean(dim=-1, keepdim=True)
    var = y.var(dim=-1, keepdim=True, unbiased=False)
    return gamma.reshape(-1) * (y - mean) / torch.sqrt(var + eps) + beta.reshape(-1)


def rmsnorm_fwd(x, w, eps=1e-5):
    """``x * rsqrt(mean(x^2)+eps) * w`` (mamba-ssm RMSNorm; eps ``modules/mamba_blocks.py:120``)."""
    return x * torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + eps) * w


def block_norm(x, sd, key, eps=1e-5):
    """The block / final norm: RMSNorm, or ``nn.LayerNorm`` when the state_dict carries a bias for it
    (``rms_norm=False``, ``modules/mamba_blocks.py:36-41,167-169``)."""
    if key + ".bias" in sd:
        if _PREC == "autocast_ref":
            raise NotImplementedError("autocast_ref models the shipped recipes (rms_norm: True)")
        return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], eps)
    return _r_ref(rmsnorm_fwd(x, sd[key + ".weight"], eps))   # the Triton RMSNorm returns the input dtype


def causal_conv_silu(xs, w, b, reverse=False):
    """Depthwise causal conv (width 4) + bias + SiLU on ``[B, L, di]``.
    ``causal_conv1d_fwd`` call site ``modules/mamba/selective_scan_interface.py:182``; the backward
    direction is the same op on the time-flipped sequence (``modules/mamba/bimamba.py:237``)."""
    x = xs.transpose(1, 2)
    if reverse:
        x = x.flip(-1)
    di, _, width = w.shape
    y = _r_any(F.silu(F.conv1d(x, w, b, padding=width - 1, groups=di)[..., : x.shape[-1]]))   # bf16 modes: u is stored bf16
    if reverse:
        y = y.flip(-1)
    return y.transpose(1, 2).contiguous()


def selective_scan(u, delta_pre, A, Bm, Cm, D, z, delta_bias, reverse=False, h_in=None,
                   impl="auto", gate=True):
    """Selective-scan recurrence, channel-last.

    u, delta_pre, z: ``[B, L, di]``; Bm, Cm: ``[B, L, Ns]``; A ``[di, Ns]``; D, delta_bias ``[di]``.
    Follows ``selective_scan_ref`` (``modules/mamba/selective_scan_interface.py:91-157``):
    ``delta = softplus(delta_pre + bias)`` (:110-112), ``h = exp(delta*A)*h + delta*B*u`` (:126-139),
    ``y = <h, C> + D*u`` (:144,:153), ``out = y * silu(z)`` (:155).  ``reverse`` runs t = L-1..0
    (the reference flips the inputs instead, ``bimamba.py:237,253``).  ``h_in``/returned
    ``h_last`` ``[B, di, Ns]`` are the chunk-carry extension needed for sequence-parallel mode
    (the reference starts from zeros, :124, and can return the last state, :147-148).

    impl: "torch" = per-step Python loop like the reference; "c" = oracle/selscan_ref.c;
    "auto" = "c" for fp32 when it builds, else "torch".
    Returns ``(out [B, L, di], h_last [B, di, Ns])``.
    """
    Bsz, L, di = u.shape
    Ns = A.shape[1]
    if impl == "auto":
        impl = "c" if u.dtype == torch.float32 else "torch"
    if impl == "c":
        lib = _clib()
        f = lambda t: np.ascontiguousarray(t.detach().to(torch.float32).numpy())
        un, dn, An, Bn, Cn, Dn, bn = f(u), f(delta_pre), f(A), f(Bm), f(Cm), f(D), f(delta_bias)
        zn = f(z) if (gate and z is not None) else None
        hn = f(h_in) if h_in is not None else None
        out = np.empty_like(un)
        hl = np.empty((Bsz, di, Ns), dtype=np.float32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
        lib.selscan_ref_f32(p(un), p(dn), p(An), p(Bn), p(Cn), p(Dn), p(zn), p(bn), p(hn), p(out),
                            Bsz, L, di, Ns, int(bool(reverse)), p(hl), None)
        return torch.from_numpy(out), torch.from_numpy(hl)
    delta = F.softplus(delta_pre + delta_bias)                           # :110-112
    h = torch.zeros(Bsz, di, Ns, dtype=u.dtype) if h_in is None else h_in.clone()
    # like the reference, materialise the whole-sequence decay and input terms first (:126, :131) ...
    deltaA = torch.exp(torch.einsum("bld,dn->bldn", delta, A))
    deltaB_u = torch.einsum("bld,bln,bld->bldn", delta, Bm, u)
    ys = [None] * L
    order = range(L - 1, -1, -1) if reverse else range(L)
    for t in order:                                                       # ... then the per-step loop (:138-151)
        h = deltaA[:, t] * h + deltaB_u[:, t]                             # :139
        ys[t] = torch.einsum("bdn,bn->bd", h, Cm[:, t])                   # :144
    y = torch.stack(ys, dim=1) + u * D                                    # :152-153
    if gate and z is not None:
        y = y * F.silu(z)                                                 # :155
    return y, h


def mixer_fwd(x, sd, prefix, scan_impl="auto", state=None):
    """Mamba mixer, ``[B,L,D]->[B,L,D]``.  Bidirectional ("v2", when the ``*_b`` parameters exist):
    ``modules/mamba/bimamba.py:176-253`` + ``MambaInnerFnNoOutProj.forward`` ``selective_scan_interface.py:164-229``.
    Unidirectional (no ``*_b`` parameters; ``mamba_ssm.Mamba`` via ``modules/mamba_blocks.py:128``): the vendored
    non-fused branch ``bimamba.py:271-310`` -- same conv / x_proj / dt_proj / scan, no 0.5 averaging.

    ``state`` (unidirectional only): dict with ``conv`` [B, di, 4] (last 4 inputs of the conv, oldest first -- the
    reference's ``conv_state`` layout, ``bimamba.py:274-277,374-380``) and ``ssm`` [B, di, Ns]; read as the history
    before this call and updated in place, which is what prefill (``:274-277,302-304``) followed by ``step``
    (``:320-372``) does one token at a time."""
    p = prefix
    W_in = sd[p + "in_proj.weight"]
    di = W_in.shape[0] // 2
    xz = _mm(x, W_in)                                                       # bimamba.py:192-196
    xs, z = xz[..., :di], xz[..., di:]                                     # ssi.py:180 (x first, z last)
    xs = _r_prod(xs)          # product bf16 mode: xz is a bf16 buffer, x half as is, z half already activated
    bidir = (p + "A_b_log") in sd
    if state is not None and bidir:
        raise ValueError("streaming state exists only for the unidirectional mixer")
    outs = []
    for sfx, rev in ((("", False), ("_b", True)) if bidir else (("", False),)):
        conv_w, conv_b = sd[p + f"conv1d{sfx}.weight"], sd[p + f"conv1d{sfx}.bias"]
        W_x, W_dt, b_dt = sd[p + f"x_proj{sfx}.weight"], sd[p + f"dt_proj{sfx}.weight"], sd[p + f"dt_proj{sfx}.bias"]
        A = -torch.exp(sd[p + ("A_log" if not rev else "A_b_log")].float())   # bimamba.py:200,222
        D = sd[p + ("D" if not rev else "D_b")].float()
        R = W_dt.shape[1]
        Ns = A.shape[1]
        h_in = None
        if state is not None:
            # history = the last 3 conv inputs (conv_state[..., 1:]); new conv_state = last 4 inputs (:274-277, :327-328)
            hist = state["conv"][:, :, 1:].transpose(1, 2).to(xs.dtype)    # [B, 3, di]
            cat = torch.cat([hist, xs], dim=1)
            u = causal_conv_silu(cat, conv_w, conv_b)[:, 3:]
            state["conv"].copy_(torch.cat([state["conv"].transpose(1, 2).to(xs.dtype), xs], dim=1)[:, -4:].transpose(1, 2))
            h_in = state["ssm"].to(xs.dtype)
        else:
            u = causal_conv_silu(xs, conv_w, conv_b, reverse=rev)         # ssi.py:182
        dbl = _mm(u, W_x)                                                  # ssi.py:186
        if _PREC == "autocast_ref":   # delta_proj_weight is cast to the autocast dtype (ssi.py:174-176), bf16 matmul
            delta_pre = _r(_r(dbl[..., :R]) @ _r(W_dt).t())
        else:
            delta_pre = dbl[..., :R] @ W_dt.t()                            # ssi.py:187 (bias NOT added here)
        Bm, Cm = dbl[..., R:R + Ns], dbl[..., R + Ns:]                     # ssi.py:193,205
        if _PREC == "fp32":
            y, h_last = selective_scan(u, delta_pre, A.to(u.dtype), Bm, Cm, D.to(u.dtype), z,
                                       b_dt.to(u.dtype), reverse=rev, impl=scan_impl, h_in=h_in)  # ssi.py:218-220
        else:
            y, h_last = selective_scan(u, delta_pre, A.to(u.dtype), Bm, Cm, D.to(u.dtype), None,
                                       b_dt.to(u.dtype), reverse=rev, impl=scan_impl, h_in=h_in, gate=False)
            if _PREC == "autocast_ref":
                y = _r(y * F.silu(z))          # the kernel gates in fp32 and returns bf16 (ssi.py:155, out_z)
            else:
                y = y * _r(F.silu(z))          # the in_proj epilogue stored silu(z) as bf16; y is rounded when stored (below)
        if state is not None:
            state["ssm"].copy_(h_last)                                     # bimamba.py:302-304, :357
        outs.append(y)
    if not bidir:
        return _mm(outs[0], sd[p + "out_proj.weight"])                    # bimamba.py:306
    if _PREC == "product_bf16":   # each direction's 0.5 * y is its own bf16 buffer; out_proj sums them along K
        return _mm(0.5 * outs[0], sd[p + "out_proj.weight"]) + _mm(0.5 * outs[1], sd[p + "out_proj.weight"])
    return _mm(_r_ref(0.5 * outs[0] + 0.5 * outs[1]), sd[p + "out_proj.weight"])  # bimamba.py:253 (bf16 add under autocast)


def mixer_step(x_t, sd, prefix, conv_state, ssm_state):
    """``Mamba.step`` (``modules/mamba/bimamba.py:320-372``) restated literally: one token ``x_t [B, D]``,
    ``conv_state [B, di, 4]`` and ``ssm_state [B, di, Ns]`` updated in place; returns ``[B, D]``."""
    p = prefix
    xz = x_t @ sd[p + "in_proj.weight"].t()                                # :323
    di = xz.shape[-1] // 2
    x, z = xz[:, :di], xz[:, di:]                                          # :324
    conv_state.copy_(torch.roll(conv_state, shifts=-1, dims=-1))           # :328
    conv_state[:, :, -1] = x                                               # :329
    x = torch.sum(conv_state * sd[p + "conv1d.weight"][:, 0, :], dim=-1) + sd[p + "conv1d.bias"]   # :330-332
    x = F.silu(x)                                                          # :333
    x_db = x @ sd[p + "x_proj.weight"].t()                                 # :343
    R = sd[p + "dt_proj.weight"].shape[1]
    Ns = ssm_state.shape[-1]
    dt, Bm, Cm = x_db[:, :R], x_db[:, R:R + Ns], x_db[:, R + Ns:]          # :344
    dt = dt @ sd[p + "dt_proj.weight"].t()                                 # :346
    A = -torch.exp(sd[p + "A_log"].float())                                # :347
    dt = F.softplus(dt + sd[p + "dt_proj.bias"])                           # :352
    dA = torch.exp(torch.einsum("bd,dn->bdn", dt, A))                      # :353
    dB = torch.einsum("bd,bn->bdn", dt, Bm)                                # :354
    ssm_state.copy_(ssm_state * dA + x.unsqueeze(-1) * dB)                 # :355
    y = torch.einsum("bdn,bn->bd", ssm_state, Cm) + sd[p + "D"] * x        # :356-357
    y = y * F.silu(z)                                                      # :358
    return y @ sd[p + "out_proj.weight"].t()                               # :365


def new_stream_state(sd, n_mamba, batch, prefix="mamba_net."):
    """Zero conv / ssm caches per layer (``allocate_inference_cache``, ``bimamba.py:368-380``)."""
    out = []
    for i in range(n_mamba):
        A = sd[f"{prefix}layers.{i}.mixer.A_log"]
        out.append({"conv": torch.zeros(batch, A.shape[0], 4), "ssm": torch.zeros(batch, A.shape[0], A.shape[1])})
    return out


def mamba_stack_fwd(h, sd, n_mamba, prefix="mamba_net.", scan_impl="auto", taps=None, states=None):
    """``MambaBlocksSequential.forward`` non-fused branch (``modules/mamba_blocks.py:186-197``)
    over ``Block.forward`` (``modules/mamba/bimamba.py:445-462``): Add -> RMSNorm -> Mixer.
    ``states`` (list from ``new_stream_state``): the ``inference_params`` caches of the reference, for streaming."""
    residual = None
    for i in range(n_mamba):
        p = f"{prefix}layers.{i}."
        residual = h if residual is None else _r_ref(h + residual)        # bimamba.py:446 (bf16 stream when residual_in_fp32 False)
        hn = block_norm(residual, sd, p + "norm")                          # bimamba.py:447
        h = mixer_fwd(hn, sd, p + "mixer.", scan_impl, state=None if states is None else states[i])  # bimamba.py:461
        if taps is not None:
            taps.append(h)
    residual = _r_ref(h + residual) if residual is not None else h         # mamba_blocks.py:196
    return block_norm(residual, sd, prefix + "norm_f")                     # mamba_blocks.py:197


def masknet_fwd(mix_w, sd, n_mamba, n_spk=2, scan_impl="auto", taps=None, mask_nonlinear="relu"):
    """``MaskNet.forward`` (``modules/mamba_masknet.py:101-139``) on channel-last ``mix_w [B,L,N]``;
    returns the mask ``[n_spk, B, L, N]`` (channel index ``s*N + n``, :126-131; ReLU :136)."""
    B, L, N = mix_w.shape
    y = cln_fwd(mix_w, sd["layer_norm.gamma"], sd["layer_norm.beta"])      # :118
    y = _mm(y, sd["bottleneck_conv1x1.conv.weight"][:, :, 0])             # :121
    y = mamba_stack_fwd(y, sd, n_mamba, scan_impl=scan_impl, taps=taps)    # :122
    score = _mm(y, sd["mask_conv1x1.conv.weight"][:, :, 0])               # :123
    score = score.reshape(B, L, n_spk, N).permute(2, 0, 1, 3)              # :126-131 ([spk, B, L, N] here)
    if mask_nonlinear == "softmax":
        return F.softmax(score, dim=-1)    # :133-134: dim=2 of the reference's [spk, B, N, L] = the N channels
    return F.relu(score)                                                   # :136


def decoder_fwd(sep_h, w_dec):
    """``conv_transpose1d(sep_h, W[N,1,K], stride K//2)`` on channel-last ``[B,L,N]`` -> ``[B,T_est]``
    (speechbrain ``dual_path.Decoder`` == ``baseline/avse2/model.py:27-37``)."""
    k = w_dec.shape[-1]
    if _PREC == "autocast_ref":   # conv_transpose1d is an autocast op: bf16 operands, bf16 result
        return _r(F.conv_transpose1d(_r(sep_h).transpose(1, 2), _r(w_dec), stride=k // 2)[:, 0, :])
    return F.conv_transpose1d(sep_h.transpose(1, 2), w_dec, stride=k // 2)[:, 0, :]


def separate(mix, sds, n_mamba, n_spk=2, scan_impl="auto", taps=None, mask_nonlinear="relu"):
    """``Separation.compute_forward`` (``Mamba-TasNet/train_wsj0mix.py:86-111``): ``[B,T] -> [B,T,n_spk]``."""
    mix_w = encoder_fwd(mix, sds["encoder"]["conv1d.weight"])              # :89
    mask = masknet_fwd(mix_w, sds["masknet"], n_mamba, n_spk, scan_impl, taps, mask_nonlinear)  # :90
    est = torch.stack([decoder_fwd(_r_ref(mix_w * mask[s]), sds["decoder"]["weight"]) for s in range(n_spk)],
                      dim=-1)                                              # :91-101
    T, T_est = mix.shape[1], est.shape[1]
    if T > T_est:                                                          # :104-109
        est = F.pad(est, (0, 0, 0, T - T_est))
    else:
        est = est[:, :T, :]
    if taps is not None:
        taps.append(mix_w)
        taps.append(mask)
    return est


# ---------------------------------------------------------------------------------------------------------------
# DPMamba: speechbrain 1.0.0 ``Dual_Path_Model`` [3P, not vendored; published algorithm] whose forward
# is restated in the vendored subclass ``Mamba-TasNet/modules/dual_path.py:56-150`` (cited below as dp.py).
def group_norm1(x, w, b, eps=1e-8):
    """``nn.GroupNorm(1, C, eps=1e-8)`` (= speechbrain ``select_norm("ln", ...)``) on channel-last ``[B, ..., C]``: one
    mean / biased variance per utterance over everything but the batch axis, per-channel affine."""
    dims = tuple(range(1, x.dim()))
    mean = x.mean(dim=dims, keepdim=True)
    var = x.var(dim=dims, keepdim=True, unbiased=False)
    return (x - mean) / torch.sqrt(var + eps) * w + b


def dp_num_chunks(L, K):
    P = K // 2
    gap = K - (P + L % K) % K                                              # Dual_Path_Model._padding
    return 2 * ((L + gap + P) // K), gap


def dp_segment(x, K):
    """``_padding`` + ``_Segmentation`` on channel-last ``[B, L, C]`` -> ``[B, S, K, C]`` (chunk s starts at padded
    position s*K/2: the even chunks tile ``padded[:-P]``, the odd ones ``padded[P:]``)."""
    B, L, C = x.shape
    P = K // 2
    S, gap = dp_num_chunks(L, K)
    padded = F.pad(x, (0, 0, P, gap + P))
    return torch.stack([padded[:, s * P: s * P + K] for s in range(S)], dim=1), gap


def dp_over_add(x, gap):
    """``_over_add`` on ``[B, S, K, C]`` -> ``[B, L, C]``: even chunks laid end to end minus the first P frames, plus odd
    chunks laid end to end minus the last P, minus the trailing gap."""
    B, S, K, C = x.shape
    P = K // 2
    even = x[:, 0::2].reshape(B, -1, C)[:, P:]
    odd = x[:, 1::2].reshape(B, -1, C)[:, :-P]
    out = even + odd
    return out[:, :-gap] if gap > 0 else out


def dp_masknet_fwd(mix_w, sd, n_dp, K, skip_around_intra, n_mamba_stack=1, n_spk=2, scan_impl="auto", skip_n_block=0):
    """``Dual_Path_Model.forward`` (dp.py:56-150 with ``skip_n_block = 0``) + ``Dual_Computation_Block.forward`` on
    channel-last ``mix_w [B, L, N]``; returns the mask ``[n_spk, B, L, N]``."""
    B, L, N = mix_w.shape
    x = group_norm1(mix_w, sd["norm.weight"], sd["norm.bias"])             # dp.py:83
    x = _mm(x, sd["conv1d.weight"][:, :, 0])                               # dp.py:88
    x, gap = dp_segment(x, K)                                              # dp.py:97  [B, S, K, D]
    S, D = x.shape[1], x.shape[3]
    residual = x                                                           # dp.py:100
    for i in range(n_dp):                                                  # dp.py:113-119
        if skip_n_block > 0 and i % skip_n_block == 0 and i != 0:          # dp.py:114-116
            x = 0.5 * x + 0.5 * residual
        p = f"dual_mdl.{i}."
        intra = mamba_stack_fwd(x.reshape(B * S, K, D), sd, n_mamba_stack, prefix=p + "intra_mdl.",
                                scan_impl=scan_impl).reshape(B, S, K, D)
        intra = group_norm1(intra, sd[p + "intra_norm.weight"], sd[p + "intra_norm.bias"])
        if skip_around_intra:
            intra = intra + x
        inter = mamba_stack_fwd(intra.transpose(1, 2).reshape(B * K, S, D), sd, n_mamba_stack, prefix=p + "inter_mdl.",
                                scan_impl=scan_impl).reshape(B, K, S, D).transpose(1, 2)
        inter = group_norm1(inter, sd[p + "inter_norm.weight"], sd[p + "inter_norm.bias"])
        x = inter + intra
    x = torch.where(x >= 0, x, sd["prelu.weight"] * x)                     # dp.py:126
    w2 = sd["conv2d.weight"][:, :, 0, 0]
    if _PREC == "product_bf16":
        # the CUDA path applies the 1x1 conv2d AFTER the overlap-add (the same linear map on half the rows, its bias then
        # counts twice: mtn_dp_overadd_prelu_fwd), so the bf16 operand of that GEMM is the overlap-added frame
        x = _mm(dp_over_add(x, gap), w2) + 2.0 * sd["conv2d.bias"]         # [B, L, spk*D]
        x = x.reshape(B, L, n_spk, D).permute(0, 2, 1, 3).reshape(B * n_spk, L, D)
    else:
        x = _mm(x, w2) + sd["conv2d.bias"]                                 # dp.py:131  [B, S, K, spk*D]
        x = x.reshape(B, S, K, n_spk, D).permute(0, 3, 1, 2, 4).reshape(B * n_spk, S, K, D)   # dp.py:137 (view B*spks)
        x = dp_over_add(x, gap)                                            # dp.py:140  [B*spk, L, D]
    o = torch.tanh(_mm(x, sd["output.0.weight"][:, :, 0]) + sd["output.0.bias"])
    g = torch.sigmoid(_mm(x, sd["output_gate.0.weight"][:, :, 0]) + sd["output_gate.0.bias"])
    x = _mm(o * g, sd["end_conv1x1.weight"][:, :, 0])                      # dp.py:141-146
    x = F.relu(x.reshape(B, n_spk, L, N))                                  # dp.py:152-154
    return x.transpose(0, 1)                                               # dp.py:157  [spk, B, L, N]


def separate_dp(mix, sds, hp, scan_impl="auto"):
    """``Separation.compute_forward`` (``train_wsj0mix.py:86-111``) with the DPMamba mask network; ``hp``: DPHParams."""
    mix_w = encoder_fwd(mix, sds["encoder"]["conv1d.weight"])
    mask = dp_masknet_fwd(mix_w, sds["masknet"], hp.n_dp, hp.chunk_size, hp.skip_around_intra, hp.n_mamba_dp // 2,
                          hp.n_spk, scan_impl, skip_n_block=getattr(hp, "skip_n_block", 0))
    est = torch.stack([decoder_fwd(mix_w * mask[s], sds["decoder"]["weight"]) for s in range(hp.n_spk)], dim=-1)
    T, T_est = mix.shape[1], est.shape[1]
    return F.pad(est, (0, 0, 0, T - T_est)) if T > T_est else est[:, :T, :]


# ---------------------------------------------------------------------------------------------------------------
# Evaluation metrics
def cal_si_snr(source, estimate):
    """SI-SNR in dB per (utterance, channel); ``source``, ``estimate`` [B, T, C].
    Restates ``cal_si_snr`` of ``baseline/avse2/utils/dnn.py:15-57`` (the in-repo statement of the speechbrain loss
    the Mamba-TasNet recipe trains and evaluates with, ``hparams/WSJ0Mix/mambatasnet_S.yaml:163``) for full-length
    utterances, in float64, with the sign flipped back (the reference returns ``-si_snr``)."""
    EPS = 1e-8
    s = source.double() - source.double().mean(dim=1, keepdim=True)          # dnn.py:29-36 (zero-mean over time)
    a = estimate.double() - estimate.double().mean(dim=1, keepdim=True)
    dot = (a * s).sum(dim=1, keepdim=True)                                   # dnn.py:45
    energy = (s ** 2).sum(dim=1, keepdim=True) + EPS                         # dnn.py:46-48
    proj = dot * s / energy                                                  # dnn.py:49
    noise = a - proj                                                         # dnn.py:51
    ratio = (proj ** 2).sum(dim=1) / ((noise ** 2).sum(dim=1) + EPS)         # dnn.py:53-55
    return 10 * torch.log10(ratio + EPS)                                     # dnn.py:56


def pit_si_snr_improvement(est, src, mix):
    """What ``save_results`` writes per utterance (``Mamba-TasNet/train_wsj0mix.py:548-558``): PIT SI-SNR of the
    estimates, of the unprocessed mixture, and their difference, for any number of speakers (``num_spks: 3`` adds
    ``s3_sig``, ``:537-538``).  The PIT wrapper [3P speechbrain ``get_si_snr_with_pitwrapper``, published behaviour] takes
    the permutation with the smallest mean loss, i.e. the largest mean SI-SNR over the speakers (first one on ties, like
    argmin over permutations in lexicographic order).  Returns (si_snr [B], si_snr_i [B], perm [B] = lexicographic rank of
    the best assignment (2 speakers: 0 direct, 1 swapped), pairs [B, n, n] = est i vs src j)."""
    import itertools
    n = est.shape[-1]
    pairs = torch.stack([torch.stack([cal_si_snr(src[..., j:j + 1], est[..., i:i + 1])[:, 0] for j in range(n)], dim=-1)
                         for i in range(n)], dim=1)                          # [B, est i, src j]
    cands = torch.stack([sum(pairs[:, i, p[i]] for i in range(n)) / n for p in itertools.permutations(range(n))], dim=1)
    best, perm = cands.max(dim=1)                                            # max returns the first maximum
    mixture = torch.stack([mix] * n, dim=-1)                                 # train_wsj0mix.py:551-553
    base = cal_si_snr(src, mixture).mean(dim=-1)
    return best, best - base, perm, pairs
