"""Weights and launch glue of the one-launch streaming push (``mtn_stream_push_fwd``, ``csrc/mtn_stream.cu``).

A streaming chunk of a few frames makes every GEMM of the causal separator a 20-row problem; the batch plan's ~100 kernel
launches are then pure latency.  ``FusedPush`` repacks the engine's weights once into the layouts that kernel reads
(fp32 vectors + per-lane ``mma.sync`` fragments, one slab per cluster rank, see ``include/mtn_b200.h``) and launches the
whole push -- encoder to overlap-add -- as one cluster kernel per call.  It works on the same carried state as the chunked
batch plan (the reference's ``conv_state`` / ``ssm_state`` caches, ``modules/mamba/bimamba.py:374-404``, plus the encoder
overlap and the decoder tail), so a stream may mix short pushes (this path) and long ones (batch plan).
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .hparams import HParams

MAX_FRAMES = 32


def eligible(hp: HParams, mode: str) -> bool:
    """What ``mtn_stream_push_fwd`` implements (everything else streams through the batch plan)."""
    return (not hp.bidirectional and mode == "fp32" and hp.enc_dim == hp.d_model and hp.d_model in (64, 128, 256, 512)
            and hp.expand == 2 and hp.d_state == 16 and hp.d_conv == 4 and hp.n_spk == 2 and hp.kernel_size == 16
            and hp.rms_norm and hp.mask_nonlinear == "relu")


def pack_fragments(W: torch.Tensor, rows: torch.Tensor, ks: torch.Tensor) -> torch.Tensor:
    """``W`` fp32 ``[n_out, K]`` -> the ``mma.sync.m16n8k16`` A-operand fragments of ``W[rows][:, ks]`` (both multiples of
    16 long), bf16 ``[tiles of 16 rows][k-steps of 16][hi | lo][lane][8]``: lane (g, t) = (lane / 4, lane % 4) holds
    a0 = (row g, k 2t..2t+1), a1 = (row g+8, same k), a2 = (row g, k 2t+8..), a3 = (row g+8, k 2t+8..)."""
    Wt = W.index_select(0, rows).index_select(1, ks)
    C, K = Wt.shape
    assert C % 16 == 0 and K % 16 == 0, (C, K)
    hi = Wt.to(torch.bfloat16)
    lo = (Wt - hi.float()).to(torch.bfloat16)

    def frag(P):   # [C, K] -> [ct, row half h, g, ks, k half q, t, e] -> [ct, ks, g, t, (q, h), e]
        T = P.reshape(C // 16, 2, 8, K // 16, 2, 4, 2)
        return T.permute(0, 3, 2, 5, 4, 1, 6).reshape(C // 16, K // 16, 32, 8)

    return torch.stack([frag(hi), frag(lo)], dim=2).contiguous()


# Measured on a B200 (tools/stream_grid.py, profiles/r02/stream/stream_grid_S_f{20,2}.jsonl; S causal, 20-frame pushes): a
# cluster of 16 CTAs takes 0.22 ms per push and 4 of them are resident at once, 8 CTAs 0.24 ms / 12 resident, 4 CTAs 0.34 ms /
# 32 resident; beyond that the clusters queue in waves.  Relative time per wave and resident clusters by cluster size:
_WAVE_TIME = {32: 0.93, 64: 1.0, 128: 1.41}
_RESIDENT = {16: 4, 8: 12, 4: 32, 2: 64}


def channels_per_cta(batch: int, d_model: int, frames: int = MAX_FRAMES) -> int:
    """d_inner channels one CTA of a stream's cluster owns (``mtn_stream_push_args.dsl``): cluster size = 2 * d_model / dsl.
    Few streams: small slices (most CTAs per stream = lowest latency).  Many streams: larger slices, so that every stream's
    cluster is resident at once instead of queueing in waves -- as far as the kernel's shared memory for ``frames`` rows
    allows.  Picks the candidate with the smallest (waves x time per wave) from the measured table above."""
    fits = lambda dsl: 0 < int(_lib.load().mtn_stream_push_smem_bytes(frames, d_model, dsl)) <= 227 * 1024
    cands = [dsl for dsl in (32, 64, 128) if fits(dsl)]
    if not cands:
        raise _lib.MtnError(f"no streaming push kernel for d_model {d_model} at {frames} frames")
    cost = lambda dsl: -(-batch // _RESIDENT[2 * d_model // dsl]) * _WAVE_TIME[dsl]
    return min(cands, key=cost)


def _pack_layers(layers, hp: HParams, dsl: int, dev):
    """Per layer: the fp32 vector blob and the fragment blob (in_proj | x_proj | out_proj slabs per cluster rank) of
    ``mtn_stream_push_args`` from ``engine.pack_layer`` dicts, for ``dsl`` channels per CTA."""
    D, di = hp.d_model, hp.d_inner
    CL = di // dsl
    f32 = lambda planes: planes.float().sum(dim=0)           # hi + lo: exact, re-splits to the same planes
    ar = lambda a, b: torch.arange(a, b, device=dev)
    vecs, frags = [], []
    for lw in layers:
        w_in, w_x = f32(lw["w_in"]), f32(lw["w_x"])       # [2di, D], [n_dbl, di] (rows dt | B | C | zero pad)
        w_out = 0.5 * f32(lw["w_out"])                    # the batch plan packs 2 * W_out for causal stacks
        assert w_x.shape[0] % 16 == 0
        vecs.append(torch.cat([lw["norm"], lw["conv_w"][0].reshape(-1), lw["conv_b"][0], lw["w_dt"][0].t().reshape(-1),
                               lw["dt_bias"][0], lw["A2"][0].reshape(-1), lw["D"][0]]))
        f_in = [pack_fragments(w_in, torch.cat([ar(dsl * r, dsl * r + dsl), ar(di + dsl * r, di + dsl * r + dsl)]), ar(0, D))
                for r in range(CL)]
        f_x = [pack_fragments(w_x, ar(0, w_x.shape[0]), ar(dsl * r, dsl * r + dsl)) for r in range(CL)]
        f_o = [pack_fragments(w_out, ar(0, D), ar(dsl * r, dsl * r + dsl)) for r in range(CL)]
        frags.append(torch.cat([t.reshape(-1) for t in (*f_in, *f_x, *f_o)]))
    return torch.stack(vecs).contiguous(), torch.stack(frags).contiguous()


class FusedStack:
    """The Mamba stack alone through the one-launch kernel (its stack-only mode): ``MambaBlocksSequential.forward(x,
    inference_params)`` for <= 32 tokens per call, i.e. the reference's decode loop (``bimamba.py:320-372`` under
    ``modules/mamba_blocks.py:186-197``) at one kernel launch per call.  The caches are the reference's own tensors, stacked:
    ``conv [n_layers, B, 4, di]`` (time-major ``conv_state``) and ``ssm [n_layers, B, di, 16]``, updated in place."""

    def __init__(self, stack):
        hp = stack.hp
        if not eligible(hp, stack.mode):
            raise _lib.MtnError("the fused streaming kernel does not implement this stack")
        self.hp, self.device, self.stack = hp, stack.device, stack
        D = hp.d_model
        with torch.cuda.device(self.device):
            z = lambda n: torch.zeros(n, dtype=torch.float32, device=self.device)
            # head blob layout of the separator kernel; only norm_f is read in stack-only mode
            self.head = torch.cat([z(16 * D), z(D), z(D), stack.norm_f, z(16 * D)]).contiguous()
        self._packed = {}    # dsl -> (layer_vec, layer_frag), packed on first use

    def packed(self, dsl: int):
        if dsl not in self._packed:
            with torch.cuda.device(self.device):
                self._packed[dsl] = _pack_layers(self.stack.layers, self.hp, dsl, self.device)
        return self._packed[dsl]

    def run(self, x: torch.Tensor, conv: torch.Tensor, ssm: torch.Tensor, dsl: int | None = None) -> torch.Tensor:
        hp = self.hp
        B, F, D = x.shape
        di = hp.d_inner
        assert tuple(conv.shape) == (hp.n_mamba, B, 4, di) and conv.is_contiguous() and conv.dtype == torch.float32
        dsl = dsl or channels_per_cta(B, D, F)
        layer_vec, layer_frag = self.packed(dsl)
        out = torch.empty_like(x)
        ops.stream_push(None, None, None, conv, ssm, None, self.head, None, None, layer_vec, layer_frag, B=B, F=F,
                        N=D, D=D, di=di, R=hp.dt_rank, n_spk=hp.n_spk, n_layers=hp.n_mamba, first=False,
                        halo_strides=(4 * di, B * 4 * di), halo_rows=4, stack_x=x, stack_out=out, dsl=dsl)
        return out


class FusedPush:
    """Packed weights of the fused push for one engine; ``run`` launches one push."""

    def __init__(self, engine):
        hp = engine.hp
        if not eligible(hp, engine.mode):
            raise _lib.MtnError("the fused streaming push does not implement this configuration")
        self.hp, self.device, self.w = hp, engine.device, engine.w
        w = engine.w
        with torch.cuda.device(self.device):
            self.head = torch.cat([w.w_enc.t().contiguous().reshape(-1), w.gamma, w.beta, w.norm_f,
                                   w.w_dec.reshape(-1)]).contiguous()
        self._packed = {}    # dsl -> (bot_frag, mask_frag, layer_vec, layer_frag), packed on first use

    def packed(self, dsl: int):
        if dsl not in self._packed:
            hp, w, dev = self.hp, self.w, self.device
            N, D = hp.enc_dim, hp.d_model
            CL = hp.d_inner // dsl
            f32 = lambda planes: planes.float().sum(dim=0)
            ar = lambda a, b: torch.arange(a, b, device=dev)
            with torch.cuda.device(dev):
                w_bot, w_mask = f32(w.w_bot), f32(w.w_mask)           # [D, N], [2N, D]
                c = dsl // 2
                bot = torch.stack([pack_fragments(w_bot, ar(c * r, c * r + c), ar(0, N)) for r in range(CL)]).contiguous()
                mask = torch.stack([pack_fragments(w_mask, ar(dsl * r, dsl * r + dsl), ar(0, D)) for r in range(CL)]).contiguous()
                self._packed[dsl] = (bot, mask) + _pack_layers(w.layers, hp, dsl, dev)
        return self._packed[dsl]

    def run(self, chunk: torch.Tensor, in_tail: torch.Tensor, first: bool, halo: torch.Tensor, h: torch.Tensor,
            ola_tail: torch.Tensor, timeline=None, dsl: int | None = None) -> torch.Tensor:
        """``chunk`` [B, 8F] (first push of a stream: [B, 8F + 8]) -> a fresh ``est`` [B, 8F, 2].  Carried state, read and
        updated in place: ``in_tail`` [B, 8] (last samples of the previous chunk), ``halo`` [n_layers, B, 3, di], ``h``
        [n_layers, 2, B, di, 16] (direction 0 is used), ``ola_tail`` [B, 2, 8].  ``dsl``: d_inner channels per CTA (None =
        ``channels_per_cta(B, d_model)``)."""
        hp = self.hp
        B, n = chunk.shape
        F = n // 8 - (1 if first else 0)
        dsl = dsl or channels_per_cta(B, hp.d_model, F)
        bot_frag, mask_frag, layer_vec, layer_frag = self.packed(dsl)
        est = torch.empty((B, 8 * F, hp.n_spk), dtype=torch.float32, device=self.device)
        ops.stream_push(chunk, in_tail, est, halo, h, ola_tail, self.head, bot_frag, mask_frag, layer_vec,
                        layer_frag, B=B, F=F, N=hp.enc_dim, D=hp.d_model, di=hp.d_inner,
                        R=hp.dt_rank, n_spk=hp.n_spk, n_layers=hp.n_mamba, first=first, timeline=timeline, dsl=dsl)
        return est
