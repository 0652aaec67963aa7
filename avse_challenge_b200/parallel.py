"""Multi-GPU drivers for the separator: batch sharding and the sequence-parallel long-form mode.

One process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU box; ``gloo`` in the CPU tests) for the
plumbing.  The reference is single-device for inference (its only parallelism is DDP in
training, ``Mamba-TasNet/train_wsj0mix.py:718``); both modes here are B200-side additions named by BASELINE.json.

``ShardedSeparator``            utterances are independent => contiguous batch slices per rank, replicated weights,
                                **no collective on the data path** (only the final gather of the estimates).
``SequenceParallelSeparator``   one long mixture, time cut into ``world x sub_chunks`` chunks.  Everything except the
                                depthwise conv and the scan is per-token.  Per layer: (1) conv halo = 3 frames of
                                ``xs`` from each neighbour; (2) scan = reduce-then-scan: a summary pass gives every
                                chunk's transfer operator ``h -> exp2(A2*sum_delta)*h + h_end``, the summaries are
                                exchanged (all-gather, or NCCL send/recv of the folded state along the rank chain),
                                ``mtn_fold_states_fwd`` composes them into the state entering each chunk, a second
                                scan pass seeded with those states writes the output.  The sub-chunks are what fills
                                148 SMs when the batch is 1 (a rank's chunk alone would occupy 8 CTAs).

The drivers contain no arithmetic: they call a *backend* (``CudaSeqBackend`` = the C-ABI kernels) and move tensors.
The CPU tests inject an oracle-backed backend to exercise exactly this host logic at world size 2.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from ._cache import LRUDict
from .hparams import HParams


# ------------------------------------------------------------------------------------------------ plumbing
def shard_slices(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced ``[start, stop)`` ranges of ``n`` items over ``world`` ranks (sizes differ by <= 1)."""
    base, rem = divmod(n, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


class Comm:
    """``torch.distributed`` when a process group exists, a no-op single rank otherwise."""

    def __init__(self, group=None):
        self.group = group
        self.on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.on else 0
        self.world = dist.get_world_size(group) if self.on else 1

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        """``[*shape]`` on every rank -> ``[world, *shape]`` (same shape required on every rank)."""
        t = t.contiguous()
        if self.world == 1:
            return t.unsqueeze(0)
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.reshape(-1), group=self.group)
        return out.view((self.world,) + tuple(t.shape))

    def send(self, t: torch.Tensor, dst: int):
        dist.send(t.contiguous(), dst, group=self.group)

    def recv(self, like: torch.Tensor, src: int) -> torch.Tensor:
        buf = torch.empty_like(like)
        dist.recv(buf, src, group=self.group)
        return buf


# ------------------------------------------------------------------------------------------------ batch sharding
class ShardedSeparator:
    """Batch-sharded separation: rank r separates utterances ``shard_slices(B, world)[r]`` with its local engine.

    ``local_forward(mix_local [b, T]) -> est_local [b, T, n_spk]`` is the rank's ``SeparatorEngine`` (or any
    callable with that contract).  ``forward`` returns the full ``[B, T, n_spk]`` on every rank."""

    def __init__(self, local_forward, n_spk: int = 2, group=None):
        self.local_forward = local_forward
        self.n_spk = n_spk
        self.comm = Comm(group)

    def my_slice(self, batch: int) -> Tuple[int, int]:
        return shard_slices(batch, self.comm.world)[self.comm.rank]

    def forward(self, mix: torch.Tensor) -> torch.Tensor:
        B, T = mix.shape
        slices = shard_slices(B, self.comm.world)
        s, e = slices[self.comm.rank]
        bmax = max(b - a for a, b in slices)
        local = mix[s:e]
        if e > s:
            est_local = self.local_forward(local)
        else:
            est_local = mix.new_zeros((0, T, self.n_spk))
        pad = torch.zeros((bmax, T, self.n_spk), dtype=est_local.dtype, device=est_local.device)
        pad[: e - s] = est_local
        allp = self.comm.all_gather(pad)                      # the only communication: result gather
        return torch.cat([allp[r, : b - a] for r, (a, b) in enumerate(slices)], dim=0)

    __call__ = forward


# ------------------------------------------------------------------------------------------------ sequence plan
@dataclass(frozen=True)
class SeqPlan:
    """Partition of ``L`` encoder frames: rank r owns frames ``ranges[r]``, cut into ``chunks[r]`` sub-chunks of
    ``Ls`` frames (the last one of a rank may be shorter: ``last_len[r]``).  Global chunk index of sub-chunk c of rank
    r is ``r * cmax + c`` (ranks with fewer sub-chunks are padded with identity operators)."""
    L: int
    world: int
    Ls: int
    cmax: int
    ranges: Tuple[Tuple[int, int], ...]
    chunks: Tuple[int, ...]
    last_len: Tuple[int, ...]


def make_seq_plan(L: int, world: int, sub_chunks: int) -> SeqPlan:
    if L < 3 * world:
        raise ValueError(f"{L} frames cannot be cut into {world} chunks of at least 3 frames (conv halo)")
    ranges = tuple(shard_slices(L, world))
    lmax = max(b - a for a, b in ranges)
    sub_chunks = max(1, min(sub_chunks, lmax))
    Ls = -(-lmax // sub_chunks)
    chunks, last = [], []
    for a, b in ranges:
        c = -(-(b - a) // Ls)
        chunks.append(c)
        last.append((b - a) - (c - 1) * Ls)
    return SeqPlan(L, world, Ls, max(chunks), ranges, tuple(chunks), tuple(last))


# ------------------------------------------------------------------------------------------------ CUDA backend
class CudaSeqBackend:
    """The per-rank kernel calls of one sequence-parallel forward (all through the C ABI, see ``ops.py``)."""

    def __init__(self, hp: HParams, sds: dict, device, mode: str = "fp32"):
        from . import _lib, ops
        from .engine import MODES, PackedWeights
        if not torch.cuda.is_available():
            raise _lib.MtnError("CudaSeqBackend needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _lib.load()
        self.ops, self._lib = ops, _lib
        self.hp, self.mode, self.device = hp, mode, torch.device(device)
        self.P = MODES[mode]["planes"]
        self.xz_dt = torch.bfloat16 if MODES[mode]["xz_bf16"] else torch.float32
        with torch.cuda.device(self.device):
            self.w = PackedWeights(hp, sds, self.device, mode)
        self.n_layers = hp.n_mamba
        self.di, self.enc_dim = hp.d_inner, hp.enc_dim
        self.tc_dt = hp.dt_rank >= 32 and mode == "fp32"    # same rule as SeparatorEngine

    # ---- chunk set-up: encoder + cLN + bottleneck on this rank's samples
    def begin(self, mix_slice: torch.Tensor, Lr: int, Ls: int, chunks: int, last_len: int):
        hp, P, dev = self.hp, self.P, self.device
        N, D, di = hp.enc_dim, hp.d_model, hp.d_inner
        nd = self.w.n_dbl
        self.Lr, self.Ls, self.C, self.last_len = Lr, Ls, chunks, last_len
        rows = chunks * Ls                                   # >= Lr; rows beyond Lr are never valid scan steps
        e = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        T_loc = mix_slice.shape[1]
        assert hp.frames(T_loc) == Lr, (T_loc, Lr)
        self.mix = e((1, (T_loc + 7) // 8 * 8), torch.float32)
        self.mix[:, :T_loc].copy_(mix_slice.to(dev, non_blocking=True))
        self.T_loc = T_loc
        self.mix_w = e((Lr, N), torch.float32)
        self.yn = e((P, Lr, N), torch.bfloat16)
        self.h = e((Lr, D), torch.float32)
        self.res = e((Lr, D), torch.float32)
        self.xn = e((P, Lr, D), torch.bfloat16)
        self.xz = e((rows, 2 * di), self.xz_dt)
        self.u = e((P, rows, 2 * di), torch.bfloat16)
        self.dbl = e((rows, 2 * nd), torch.float32)
        self.dtp = e((rows, 2, 2, self.ops.rp_for(hp.dt_rank)), torch.bfloat16)
        self.y = e((P, rows, 2 * di), torch.bfloat16)
        self.sep_full = e((Lr + 1, hp.n_spk * N), torch.float32)   # row 0 = last frame of the previous rank
        o, w = self.ops, self.w
        o.encoder_cln(self.mix, w.w_enc, w.gamma, w.beta, P, mix_w=self.mix_w, yn=self.yn, T=T_loc)
        o.gemm(self.yn, w.w_bot, Lr, D, N, out=self.h)

    # ---- per layer
    def pre(self, i: int):
        lw, o, hp = self.w.layers[i], self.ops, self.hp
        o.add_rmsnorm(self.h, self.res, i > 0, lw["norm"], self.P, xn=self.xn, beta=lw["norm_b"])
        o.gemm(self.xn, lw["w_in"], self.Lr, 2 * hp.d_inner, hp.d_model, out=self.xz, epilogue=self._lib.EPI_INPROJ,
               epi_param=hp.d_inner, out_bf16=self.xz.dtype == torch.bfloat16)

    def xs_edges(self) -> torch.Tensor:
        """First and last three ``xs`` rows of this rank's chunk, fp32 ``[2, 3, di]`` (what the neighbours need)."""
        di = self.di
        return torch.stack([self.xz[:3, :di], self.xz[self.Lr - 3:self.Lr, :di]]).float().contiguous()

    def conv_xproj(self, i: int, halo_lo: Optional[torch.Tensor], halo_hi: Optional[torch.Tensor]):
        lw, o, hp = self.w.layers[i], self.ops, self.hp
        di, nd = hp.d_inner, self.w.n_dbl
        lo = halo_lo.reshape(1, 3, di).contiguous() if halo_lo is not None else None
        hi = halo_hi.reshape(1, 3, di).contiguous() if halo_hi is not None else None
        o.conv_silu(self.xz, lw["conv_w"], lw["conv_b"], 1, self.Lr, di, self.P, u=self.u, halo_lo=lo, halo_hi=hi)
        if self.tc_dt:
            o.gemm(self.u, lw["w_x"], self.Lr, nd, di, out=self.dbl, groups=2, out_group_stride=nd,
                   epilogue=self._lib.EPI_XPROJ, epi_param=o.rp_for(hp.dt_rank), aux=self.dtp)
        else:
            o.gemm(self.u, lw["w_x"], self.Lr, nd, di, out=self.dbl, groups=2, out_group_stride=nd)

    def _scan(self, i: int, **kw):
        lw, hp = self.w.layers[i], self.hp
        return self.ops.scan(self.u, self.dbl, self.xz, hp.d_inner, lw["w_dt"], lw["dt_bias"], lw["A2"], lw["D"], self.C,
                             self.Ls, hp.d_inner, hp.dt_rank, L_last=self.last_len, dtp=self.dtp if self.tc_dt else None, **kw)

    def scan_summary(self, i: int):
        """Summary pass: ``(h_end [2, C, di, 16], sum_delta [2, C, di])`` of this rank's sub-chunks, h_in = 0."""
        di = self.di
        h_end = torch.zeros((2, self.C, di, 16), dtype=torch.float32, device=self.device)
        sdl = torch.zeros((2, self.C, di), dtype=torch.float32, device=self.device)
        self._scan(i, h_out=h_end, sum_delta=sdl, summary_only=True)
        return h_end, sdl

    def fold(self, i: int, h_end: torch.Tensor, sdl: torch.Tensor, g0: int, n_out: int, h0=None, want_final=False,
             dir_mask: int = 3):
        return self.ops.fold_states(h_end.contiguous(), sdl.contiguous(), self.w.layers[i]["A2"], g0, n_out, h0=h0,
                                    want_final=want_final, dir_mask=dir_mask)

    def scan_seeded(self, i: int, h_in: torch.Tensor):
        self._scan(i, y=self.y, h_in=h_in.contiguous())

    def out_proj(self, i: int):
        lw, hp = self.w.layers[i], self.hp
        self.ops.gemm(self.y, lw["w_out"], self.Lr, hp.d_model, 2 * hp.d_inner, out=self.h)

    # ---- tail
    def head(self):
        hp, w, o = self.hp, self.w, self.ops
        o.add_rmsnorm(self.h, self.res, True, w.norm_f, self.P, xn=self.xn, beta=w.norm_f_b)
        o.gemm(self.xn, w.w_mask, self.Lr, hp.n_spk * hp.enc_dim, hp.d_model, out=self.sep_full[1:],
               epilogue=self._lib.EPI_MASK, epi_param=hp.enc_dim, aux=self.mix_w)

    def sep_last_row(self) -> torch.Tensor:
        return self.sep_full[self.Lr].clone()

    def set_sep_halo(self, row: Optional[torch.Tensor]):
        if row is None:
            self.sep_full[0].zero_()
        else:
            self.sep_full[0].copy_(row)

    def decode(self) -> torch.Tensor:
        """Overlap-add decode of frames ``-1 .. Lr-1`` (frame -1 = the previous rank's last frame): ``[8*(Lr+2), n_spk]``
        whose sample 0 is sample ``8*(f0 - 1)`` of the recording."""
        hp = self.hp
        Lx = self.Lr + 1
        T_x = (Lx - 1) * 8 + 16
        est = self.ops.decoder(self.sep_full, self.w.w_dec, 1, T_x, Lx, hp.enc_dim, hp.n_spk)
        return est[0]


# ------------------------------------------------------------------------------------------------ driver
class SequenceParallelSeparator:
    """``forward(mix [1, T]) -> est [1, T, n_spk]`` with time sharded over the ranks of ``group``.

    ``mix`` must be the same (replicated) tensor on every rank; the result is assembled on every rank.
    ``exchange``: ``"allgather"`` (default: one all-gather of the chunk summaries per layer) or ``"sendrecv"``
    (the folded state is handed down the rank chain with point-to-point send/recv; same numbers)."""

    # sub_chunks = 74: every chunk is 2 * d_inner / 32 independent warp pairs for the scan (32 for S / M, 64 for L) and a
    # B200 holds 148 SMs x 8 pairs = 1 184 of them, so 74 chunks are exactly 2 (S) or 4 (L) full waves; 64 chunks left the
    # second wave 73 % full (config 5 on one GPU: 361 ms at 64, 341 ms at 74, 340 ms at 128-256)
    def __init__(self, hp: HParams, sds: Optional[dict] = None, device="cuda", mode: str = "fp32", sub_chunks: int = 74,
                 exchange: str = "allgather", group=None, backend=None, use_graph: bool = True):
        if exchange not in ("allgather", "sendrecv"):
            raise ValueError("exchange must be 'allgather' or 'sendrecv'")
        if not hp.bidirectional:
            raise NotImplementedError("the chunked-scan plan is built for the bidirectional stack; a causal model "
                                      "streams through StreamingSeparator instead")
        self.hp, self.sub_chunks, self.exchange = hp, sub_chunks, exchange
        self.comm = Comm(group)
        self.be = backend if backend is not None else CudaSeqBackend(hp, sds, device, mode)
        # On a single rank there is no collective in the path, so the whole chunked forward replays as one CUDA graph
        # per length.  That is also the low-latency plan for ONE short utterance: the scan's serial chain is sub_chunks
        # times shorter than in the batch plan (4 s @ 8 kHz, S: 9.0 ms -> 2.1 ms at 16 sub-chunks).
        self.use_graph = use_graph and backend is None
        self._graphs = LRUDict()   # one whole-forward graph (with its private buffers) per recording length, LRU-bounded

    @torch.no_grad()
    def forward(self, mix: torch.Tensor) -> torch.Tensor:
        if mix.dim() != 2 or mix.shape[0] != 1:
            raise ValueError("sequence-parallel mode separates one recording: mix must be [1, T]")
        if not (self.use_graph and self.comm.world == 1 and mix.is_cuda):
            return self._forward(mix)
        T = mix.shape[1]
        ent = self._graphs.get(T)
        if ent is None:
            static_in = mix.clone()
            self._forward(static_in)                     # eager first: workspaces, kernel attributes, shape checks
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = self._forward(static_in)
            ent = self._graphs[T] = (g, static_in, static_out)
        g, static_in, static_out = ent
        static_in.copy_(mix, non_blocking=True)
        g.replay()
        return static_out.clone()

    def _forward(self, mix: torch.Tensor) -> torch.Tensor:
        hp, be, comm = self.hp, self.be, self.comm
        r, W = comm.rank, comm.world
        T = mix.shape[1]
        L = hp.frames(T)
        plan = make_seq_plan(L, W, self.sub_chunks)
        f0, f1 = plan.ranges[r]
        Lr, C, cmax = f1 - f0, plan.chunks[r], plan.cmax
        be.begin(mix[:, 8 * f0: 8 * f1 + 8], Lr, plan.Ls, C, plan.last_len[r])
        for i in range(be.n_layers):
            be.pre(i)
            edges = comm.all_gather(be.xs_edges())                           # [W, 2, 3, di]
            be.conv_xproj(i, edges[r - 1, 1] if r > 0 else None, edges[r + 1, 0] if r < W - 1 else None)
            h_end, sdl = be.scan_summary(i)                                  # [2, C, di, 16], [2, C, di]
            if self.exchange == "allgather":
                he = h_end.new_zeros((2, cmax) + tuple(h_end.shape[2:]))
                sd = sdl.new_zeros((2, cmax) + tuple(sdl.shape[2:]))
                he[:, :C], sd[:, :C] = h_end, sdl                            # padding chunks = identity operators
                he_all = comm.all_gather(he).transpose(0, 1).reshape((2, W * cmax) + tuple(h_end.shape[2:]))
                sd_all = comm.all_gather(sd).transpose(0, 1).reshape((2, W * cmax) + tuple(sdl.shape[2:]))
                h_in, _ = be.fold(i, he_all, sd_all, r * cmax, C)
            else:
                h_in = self._fold_chain(i, h_end, sdl, C)
            be.scan_seeded(i, h_in)
            be.out_proj(i)
        be.head()
        rows = comm.all_gather(be.sep_last_row())                            # [W, n_spk*N]
        be.set_sep_halo(rows[r - 1] if r > 0 else None)
        est_x = be.decode()                                                  # samples 8*(f0-1) .. 8*(f1+1)
        n_keep = 8 * Lr + (8 if r == W - 1 else 0)
        piece = est_x[8: 8 + n_keep]
        nmax = 8 * max(b - a for a, b in plan.ranges) + 8
        pad = piece.new_zeros((nmax, piece.shape[1]))
        pad[:n_keep] = piece
        allp = comm.all_gather(pad)                                          # result gather
        parts = [allp[q, : 8 * (b - a) + (8 if q == W - 1 else 0)] for q, (a, b) in enumerate(plan.ranges)]
        est = torch.cat(parts, dim=0)                                        # T_est = 8*L + 8 samples
        out = est.new_zeros((T, est.shape[1]))                               # pad / trim (train_wsj0mix.py:104-109)
        n = min(T, est.shape[0])
        out[:n] = est[:n]
        return out.unsqueeze(0)

    __call__ = forward

    def _fold_chain(self, i, h_end, sdl, C):
        """Point-to-point variant: rank r receives the state entering its chunk from its neighbour, folds its own
        sub-chunks and passes the state leaving its chunk on.  Forward chain 0 -> W-1, backward chain W-1 -> 0."""
        be, comm = self.be, self.comm
        r, W = comm.rank, comm.world
        like = h_end[:, 0].contiguous()                                      # [2, di, 16] (both directions' slots)
        h0 = torch.zeros_like(like)
        if r > 0:
            h0[0] = comm.recv(like[0], r - 1)
        hin_f, fin_f = be.fold(i, h_end, sdl, 0, C, h0=h0, want_final=True, dir_mask=1)
        if r < W - 1:
            comm.send(fin_f[0], r + 1)
        if r < W - 1:
            h0[1] = comm.recv(like[1], r + 1)
        hin_b, fin_b = be.fold(i, h_end, sdl, 0, C, h0=h0, want_final=True, dir_mask=2)
        if r > 0:
            comm.send(fin_b[1], r - 1)
        h_in = hin_f.clone()
        h_in[1] = hin_b[1]
        return h_in
