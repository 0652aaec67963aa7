"""DPMamba (dual-path Mamba) separator on one B200 -- SURVEY.md 8f rank 1.

The ``dpmamba_*`` recipes (``Mamba-TasNet/hparams/WSJ0Mix/dpmamba_{XS,S,M,L}.yaml``) keep the Encoder / Decoder and
``compute_forward`` of Mamba-TasNet and replace the mask network by speechbrain's ``Dual_Path_Model`` [third party; forward
restated in the vendored ``Mamba-TasNet/modules/dual_path.py:56-150``] with one-layer ``MambaBlocksSequential`` stacks as
intra- and inter-chunk models: the frame sequence is cut into chunks of K = 250 frames at 50 % overlap, and every dual
block runs a bidirectional Mamba along each chunk (``B*S`` sequences of K frames), a whole-utterance GroupNorm, then a
Mamba across chunks (``B*K`` sequences of S frames) and another GroupNorm.

Plan (everything channel-last; the 4-D tensor [B, N, K, S] of the reference is rows (b, s, k) x D):

    encoder -> GroupNorm(1, N) -> 1x1 conv N->D (GEMM) -> segmentation (row gather, zero padded)
    n_dp x [ intra stack (6 + 1 launches, the Mamba-TasNet kernels) -> GN stats -> GN apply + skip, also written
             transposed to rows (b, k, s) -> inter stack -> GN stats -> GN apply (transposed read) + residual ]
    PReLU + overlap-add (row gather-add) -> conv2d 1x1 D -> spk*D (GEMM; moved behind the overlap-add: same linear
    map on half the rows, bias counted twice) -> bias -> [output | gate] GEMM per speaker group -> tanh * sigmoid ->
    end_conv1x1 GEMM with the relu * mix_w epilogue -> decoder.

The two stacks of a block share one activation workspace (B*S*K == B*K*S rows).  No CPU / eager fallback.
"""
from __future__ import annotations

import copy

import torch

from . import _lib, ops
from ._cache import LRUDict
from .engine import MODES, LayerWorkspace, MambaStack
from .hparams import DPHParams


class DPWorkspace:
    def __init__(self, hp: DPHParams, batch: int, T: int, device, mode: str):
        P = MODES[mode]["planes"]
        N, D, K, spk = hp.enc_dim, hp.d_model, hp.chunk_size, hp.n_spk
        L = hp.frames(T)
        S = ops.dp_num_chunks(L, K)
        M, M2 = batch * L, batch * S * K
        e = lambda shape, dt=torch.float32: torch.empty(shape, dtype=dt, device=device)
        self.batch, self.T, self.L, self.S, self.K, self.M, self.M2 = batch, T, L, S, K, M, M2
        self.mix = torch.zeros((batch, (T + 7) // 8 * 8), dtype=torch.float32, device=device)
        self.mix_w = e((M, N))
        self.yn = e((P, M, N), torch.bfloat16)          # encoder scratch, then the GroupNorm'ed frames (GEMM operand)
        self.xc = e((M, D))
        self.X = e((M2, D))                             # dual-path tensor, rows (b, s, k)
        self.I = e((M2, D))                             # intra branch output (kept for the block's residual)
        self.O = e((M2, D))                             # stack output before its GroupNorm
        self.R = e((M2, D)) if hp.skip_n_block > 0 else None   # segmented input kept for Dual_Path_Model_Skip's blend
        self.partials = ops.gn_partials(batch, max(S * K, L), max(N, D), device)
        self.intra = LayerWorkspace(hp.stack, batch * S, K, device, mode)
        self.inter = copy.copy(self.intra)              # same buffers, viewed as B*K sequences of S frames
        self.inter.batch, self.inter.L = batch * K, S
        self.Yp = e((P, M, D), torch.bfloat16)
        self.c2 = e((M, spk * D))
        self.c2p = e((P, M, spk * D), torch.bfloat16)
        self.og = e((M, spk * 2 * D))
        self.gp = e((P, M, spk * D), torch.bfloat16)
        self.sep = e((M, spk * N))
        self.frames = e((M, spk, 16))
        self.est = e((batch, T, spk))

    def nbytes(self):
        own = sum(t.numel() * t.element_size() for t in vars(self).values() if isinstance(t, torch.Tensor))
        return own + self.intra.nbytes()


class DPSeparatorEngine:
    """mix [B, T] fp32 (CUDA) -> est_source [B, T, n_spk] fp32 with the DPMamba mask network."""

    def __init__(self, hp: DPHParams, sds: dict, device="cuda", mode: str = "fp32", use_graph: bool = True):
        if mode not in MODES:
            raise ValueError(f"mode must be one of {list(MODES)}")
        if not torch.cuda.is_available():
            raise _lib.MtnError("DPSeparatorEngine needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if hp.n_spk < 1 or hp.chunk_size % 2 or hp.n_mamba_dp < 2:
            raise NotImplementedError("DPMamba: n_spk >= 1, chunk_size even, n_mamba_dp >= 2")
        _lib.load()
        self.hp, self.mode, self.device, self.use_graph = hp, mode, torch.device(device), use_graph
        P = self.P = MODES[mode]["planes"]
        N, D, spk = hp.enc_dim, hp.d_model, hp.n_spk
        m = sds["masknet"]
        f32 = lambda t: t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(self.device):
            self.w_enc = f32(sds["encoder"]["conv1d.weight"]).reshape(N, hp.kernel_size)
            self.w_dec = f32(sds["decoder"]["weight"]).reshape(N, hp.kernel_size)
            self.ones, self.zeros = torch.ones(N, device=self.device), torch.zeros(N, device=self.device)
            self.norm_w, self.norm_b = f32(m["norm.weight"]), f32(m["norm.bias"])
            self.w_conv1d = ops.split_planes(f32(m["conv1d.weight"]).reshape(D, N), P)
            self.blocks = []
            for i in range(hp.n_dp):
                p = f"dual_mdl.{i}."
                self.blocks.append({
                    "intra": MambaStack(hp.stack, m, device=self.device, mode=mode, prefix=p + "intra_mdl."),
                    "inter": MambaStack(hp.stack, m, device=self.device, mode=mode, prefix=p + "inter_mdl."),
                    "intra_w": f32(m[p + "intra_norm.weight"]), "intra_b": f32(m[p + "intra_norm.bias"]),
                    "inter_w": f32(m[p + "inter_norm.weight"]), "inter_b": f32(m[p + "inter_norm.bias"]),
                })
            self.prelu_w = f32(m["prelu.weight"]).reshape(1)
            self.w_conv2d = ops.split_planes(f32(m["conv2d.weight"]).reshape(spk * D, D), P)
            self.b_conv2d = f32(m["conv2d.bias"])
            wo, wg = f32(m["output.0.weight"]).reshape(D, D), f32(m["output_gate.0.weight"]).reshape(D, D)
            self.w_og = ops.split_planes(torch.cat([wo, wg] * spk, dim=0).contiguous(), P)       # [P, spk*2D, D]
            self.b_o, self.b_g = f32(m["output.0.bias"]), f32(m["output_gate.0.bias"])
            self.w_end = ops.split_planes(f32(m["end_conv1x1.weight"]).reshape(N, D).repeat(spk, 1).contiguous(), P)
        # LRU-bounded like SeparatorEngine's (engine.LRUDict): evicting a workspace drops the graph captured against it
        self._graphs = LRUDict()
        self._ws = LRUDict(on_evict=lambda key, ws: self._graphs.pop(key, None))
        self._prof = None
        # encoder, gn x2, conv1d, segment | per block: 2 x (stack 7 + gn 2) | overadd, conv2d, bias, og, gate, end, decoder(2)
        self.launches_per_forward = 5 + hp.n_dp * 2 * (6 * (hp.n_mamba_dp // 2) + 1 + 2) + 8

    def workspace(self, batch, T) -> DPWorkspace:
        key = (batch, T)
        if key not in self._ws:
            self._ws[key] = DPWorkspace(self.hp, batch, T, self.device, self.mode)
        return self._ws[key]

    def _op(self, name, fn, *a, **k):
        prof = self._prof
        if prof is None:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        prof.append((name, e0, e1))
        return out

    def _run(self, ws: DPWorkspace, mask_only: bool = False):
        """``mask_only``: start from ``ws.mix_w`` (already filled) and stop at the ReLU mask in ``ws.sep`` -- the
        stand-alone ``Dual_Path_Model.forward``."""
        hp, P, op = self.hp, self.P, self._op
        N, D, K, S, B, L, M, spk = hp.enc_dim, hp.d_model, ws.K, ws.S, ws.batch, ws.L, ws.M, hp.n_spk
        if not mask_only:
            op("encoder", ops.encoder_cln, ws.mix, self.w_enc, self.ones, self.zeros, P, mix_w=ws.mix_w, yn=ws.yn, T=ws.T)
        op("gn_stats", ops.gn_stats, ws.mix_w, B, L, N, ws.partials)                       # dual_path.py:83
        op("gn_apply", ops.gn_apply, ws.mix_w, ws.partials, self.norm_w, self.norm_b, B, 1, L, N, planes=ws.yn)
        op("gemm_conv1d", ops.gemm, ws.yn, self.w_conv1d, M, D, N, out=ws.xc)               # dual_path.py:88
        op("dp_segment", ops.dp_segment, ws.xc, B, L, D, K, S, ws.X, ws.intra.h)            # dual_path.py:97
        n_skip = hp.skip_n_block
        if n_skip > 0:
            ws.R.copy_(ws.X)                                                               # residual = x, dual_path.py:100
        fused = D in (128, 256, 512) and hp.stack.rms_norm   # GroupNorm apply + the next stack's opening RMSNorm in one kernel (gn_apply_norm)
        for i, blk in enumerate(self.blocks):
            last = i == len(self.blocks) - 1
            # `x = 0.5 * x + 0.5 * residual` in front of block i + 1 (dual_path.py:114-116) is applied by the norm that ends
            # block i
            blend = ws.R if (n_skip > 0 and not last and (i + 1) % n_skip == 0) else None
            blk["intra"].run(ws.intra, ws.O, prenormed=fused and i > 0)                     # rows (b, s, k)
            op("gn_stats", ops.gn_stats, ws.O, B, S * K, D, ws.partials)
            skip = ws.X if hp.skip_around_intra else None
            if fused:
                op("gn_apply_norm", ops.gn_apply_norm, ws.O, ws.partials, blk["intra_w"], blk["intra_b"], B, S, K, D,
                   skip=skip, out_a=ws.I, res_next=ws.inter.res, xn_next=ws.inter.xn,
                   g_next=blk["inter"].layers[0]["norm"], next_transposed=True)
            else:
                op("gn_apply", ops.gn_apply, ws.O, ws.partials, blk["intra_w"], blk["intra_b"], B, S, K, D, skip=skip,
                   out_a=ws.I, out_t=ws.inter.h)
            blk["inter"].run(ws.inter, ws.O, prenormed=fused)                               # rows (b, k, s)
            op("gn_stats", ops.gn_stats, ws.O, B, S * K, D, ws.partials)
            if fused and not last:
                op("gn_apply_norm", ops.gn_apply_norm, ws.O, ws.partials, blk["inter_w"], blk["inter_b"], B, S, K, D,
                   skip=ws.I, out_a=ws.X, x_transposed=True, blend=blend, res_next=ws.intra.res, xn_next=ws.intra.xn,
                   g_next=self.blocks[i + 1]["intra"].layers[0]["norm"], next_transposed=False)
            else:
                op("gn_apply", ops.gn_apply, ws.O, ws.partials, blk["inter_w"], blk["inter_b"], B, S, K, D, skip=ws.I,
                   out_a=ws.X, out_a2=None if last else ws.intra.h, x_transposed=True, blend=blend)
        op("dp_overadd_prelu", ops.dp_overadd_prelu, ws.X, self.prelu_w, ws.Yp, B, L, D, K, S)   # dual_path.py:126,140
        op("gemm_conv2d", ops.gemm, ws.Yp, self.w_conv2d, M, spk * D, D, out=ws.c2)         # dual_path.py:131
        op("bias_planes", ops.bias_planes, ws.c2, self.b_conv2d, 2.0, ws.c2p, M, spk * D)
        op("gemm_out_gate", ops.gemm, ws.c2p, self.w_og, M, 2 * D, D, out=ws.og, groups=spk, out_group_stride=2 * D)
        op("gate_planes", ops.gate_planes, ws.og, self.b_o, self.b_g, ws.gp, M, spk, D)     # dual_path.py:141
        if mask_only:
            op("gemm_end_relu", ops.gemm, ws.gp, self.w_end, M, N, D, out=ws.sep, groups=spk, out_group_stride=N,
               epilogue=_lib.EPI_RELU)                                                      # dual_path.py:146,154
            return ws.sep
        op("gemm_end_mask", ops.gemm, ws.gp, self.w_end, M, N, D, out=ws.sep, groups=spk, out_group_stride=N,
           epilogue=_lib.EPI_MASK, epi_param=N, aux=ws.mix_w)                               # :146,:154 + train_wsj0mix.py:91-92
        op("decoder", ops.decoder, ws.sep, self.w_dec, B, ws.T, L, N, spk, est=ws.est, frames=ws.frames)
        return ws.est

    def _set_prof(self, prof):
        self._prof = prof
        for blk in self.blocks:
            blk["intra"]._prof = prof
            blk["inter"]._prof = prof

    def profile_ops(self, batch: int, T: int, steps: int = 1):
        ws = self.workspace(batch, T)
        agg = {}
        for _ in range(steps):
            self._set_prof([])
            try:
                self._run(ws)
                torch.cuda.current_stream().synchronize()
                for name, e0, e1 in self._prof:
                    a = agg.setdefault(name, [0, 0.0])
                    a[0] += 1
                    a[1] += e0.elapsed_time(e1)
            finally:
                self._set_prof(None)
        return {k: {"launches": n // steps, "ms": t / n, "ms_per_forward": t / steps} for k, (n, t) in agg.items()}

    def forward_into_workspace(self, batch: int, T: int):
        ws = self.workspace(batch, T)
        if not self.use_graph:
            return self._run(ws)
        key = (batch, T)
        g = self._graphs.get(key)
        if g is None:
            self._run(ws)
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run(ws)
            self._graphs[key] = g
        g.replay()
        return ws.est

    @torch.no_grad()
    def forward(self, mix: torch.Tensor) -> torch.Tensor:
        if mix.dim() != 2 or mix.dtype != torch.float32 or not mix.is_cuda:
            raise _lib.MtnError("forward expects a CUDA fp32 tensor of shape [batch, T]")
        B, T = mix.shape
        if T < 16:
            raise _lib.MtnError(f"T={T}: need at least one 16-sample frame")
        ws = self.workspace(B, T)
        ws.mix[:, :T].copy_(mix, non_blocking=True)
        return self.forward_into_workspace(B, T).clone()

    __call__ = forward
