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
                                exchanged (ONE all-gather of a packed record per layer, or NCCL send/recv of the folded
                                state along the rank chain),
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
        self.on = group != "local" and dist.is_available() and dist.is_initialized()   # "local": this process alone
        self.rank = dist.get_rank(group) if self.on else 0
        self.world = dist.get_world_size(group) if self.on else 1

    def all_gather(self, t: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``[*shape]`` on every rank -> ``[world, *shape]`` (same shape required on every rank).  ``out``: a
        pre-allocated contiguous ``[world, *shape]`` receive buffer (no allocation on the hot path)."""
        t = t.contiguous()
        if self.world == 1:
            return t.unsqueeze(0)
        if out is None:
            out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out.view(-1), t.reshape(-1), group=self.group)
        return out

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
    """The per-rank kernel calls of one sequence-parallel forward (all through the C ABI, see ``ops.py``).

    All activation buffers of a (local length, plan) shape are allocated once and reused (``_Shape``, LRU-bounded), so a
    forward allocates nothing -- which is also what makes it capturable in a CUDA graph together with its collectives."""

    def __init__(self, hp: HParams, sds: Optional[dict], device, mode: str = "fp32", weights=None):
        """``weights``: an already packed ``engine.PackedWeights`` of the same hparams / mode / device (shares the device
        copies with a ``SeparatorEngine``); otherwise packed here from ``sds``."""
        from . import _lib, ops
        from .engine import MODES, PackedWeights, resolve_device
        if not torch.cuda.is_available():
            raise _lib.MtnError("CudaSeqBackend needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _lib.load()
        self.ops, self._lib = ops, _lib
        self.hp, self.mode, self.device = hp, mode, resolve_device(device)
        self.P = MODES[mode]["planes"]
        self.xz_dt = torch.bfloat16 if MODES[mode]["xz_bf16"] else torch.float32
        if weights is not None:
            self.w = weights
        else:
            with torch.cuda.device(self.device):
                self.w = PackedWeights(hp, sds, self.device, mode)
        self.n_layers = hp.n_mamba
        self.di, self.enc_dim = hp.d_inner, hp.enc_dim
        self.tc_dt = hp.dt_rank >= 32 and mode == "fp32"    # same rule as SeparatorEngine
        self.cmax, self.world = None, 1
        self._shapes = LRUDict()

    def set_plan(self, cmax: int, world: int):
        """Sub-chunks per rank (padded) and the world size: fixes the layout of the packed summary record."""
        self.cmax, self.world = cmax, world

    # ---- chunk set-up: encoder + cLN + bottleneck on this rank's samples
    def begin(self, mix_slice: torch.Tensor, Lr: int, Ls: int, chunks: int, last_len: int):
        hp, P, dev = self.hp, self.P, self.device
        N, D, di = hp.enc_dim, hp.d_model, hp.d_inner
        nd = self.w.n_dbl
        T_loc = mix_slice.shape[1]
        assert hp.frames(T_loc) == Lr, (T_loc, Lr)
        cmax = self.cmax if self.cmax is not None else chunks
        key = (T_loc, Lr, Ls, chunks, last_len, cmax, self.world)
        ws = self._shapes.get(key)
        if ws is None:
            rows = chunks * Ls                                   # >= Lr; rows beyond Lr are never valid scan steps
            z = lambda shape, dt=torch.float32: torch.zeros(shape, dtype=dt, device=dev)
            ws = {}
            ws["mix"] = z((1, (T_loc + 7) // 8 * 8))
            ws["mix_w"] = z((Lr, N))
            ws["yn"] = z((P, Lr, N), torch.bfloat16)
            ws["h"], ws["res"] = z((Lr, D)), z((Lr, D))
            ws["xn"] = z((P, Lr, D), torch.bfloat16)
            ws["xz"] = z((rows, 2 * di), self.xz_dt)
            ws["u"] = z((P, rows, 2 * di), torch.bfloat16)
            ws["dbl"] = z((rows, 2 * nd))
            ws["dtp"] = z((rows, 2, 2, self.ops.rp_for(hp.dt_rank)), torch.bfloat16)
            ws["y"] = z((P, rows, 2 * di), torch.bfloat16)
            ws["sep_full"] = z((Lr + 1, hp.n_spk * N))         # row 0 = last frame of the previous rank
            ws["edges"] = z((2, 3, di))
            ws["edges_all"] = z((self.world, 2, 3, di))
            # packed chunk summaries: [h_end (2, cmax, di, 16) | sum_delta (2, cmax, di)]; entries past `chunks` stay zero
            # (identity operators).  The scan writes [2][batch][di][..] with batch = chunks, so it can write the record
            # in place only when chunks == cmax; otherwise it writes compact buffers that are copied in.
            rec = 2 * cmax * di * 17
            ws["pack"] = z((rec,))
            ws["pack_all"] = z((self.world, rec))
            ws["h_end"] = ws["pack"][: 2 * cmax * di * 16].view(2, cmax, di, 16)
            ws["sdl"] = ws["pack"][2 * cmax * di * 16:].view(2, cmax, di)
            if chunks != cmax:
                ws["h_end_c"], ws["sdl_c"] = z((2, chunks, di, 16)), z((2, chunks, di))
            ws["h_in"] = z((2, chunks, di, 16))
            ws["row"] = z((hp.n_spk * N,))
            ws["rows_all"] = z((self.world, hp.n_spk * N))
            Lx = Lr + 1
            ws["frames"] = z((Lx, hp.n_spk, 16))
            ws["est"] = z((1, (Lx - 1) * 8 + 16, hp.n_spk))
            self._shapes[key] = ws
        self.ws = ws
        self.Lr, self.Ls, self.C, self.last_len, self.T_loc = Lr, Ls, chunks, last_len, T_loc
        for k in ("mix", "mix_w", "yn", "h", "res", "xn", "xz", "u", "dbl", "dtp", "y", "sep_full"):
            setattr(self, k, ws[k])
        self.mix[:, :T_loc].copy_(mix_slice, non_blocking=True)
        o, w = self.ops, self.w
        o.encoder_cln(self.mix, w.w_enc, w.gamma, w.beta, P, mix_w=self.mix_w, yn=self.yn, T=T_loc)
        o.gemm(self.yn, w.w_bot, Lr, D, N, out=self.res)   # the residual stream starts here; out_proj adds to it in place

    def gather_buffer(self, name: str) -> torch.Tensor:
        """Pre-allocated ``[world, ...]`` receive buffer of the collective called ``name``."""
        return self.ws[name]

    # ---- per layer
    def pre(self, i: int):
        lw, o, hp = self.w.layers[i], self.ops, self.hp
        o.add_rmsnorm(None, self.res, True, lw["norm"], self.P, xn=self.xn, beta=lw["norm_b"])
        o.gemm(self.xn, lw["w_in"], self.Lr, 2 * hp.d_inner, hp.d_model, out=self.xz, epilogue=self._lib.EPI_INPROJ,
               epi_param=hp.d_inner, out_bf16=self.xz.dtype == torch.bfloat16)

    def xs_edges(self) -> torch.Tensor:
        """First and last three ``xs`` rows of this rank's chunk, fp32 ``[2, 3, di]`` (what the neighbours need)."""
        di, e = self.di, self.ws["edges"]
        e[0].copy_(self.xz[:3, :di])
        e[1].copy_(self.xz[self.Lr - 3:self.Lr, :di])
        return e

    def conv_xproj(self, i: int, halo_lo: Optional[torch.Tensor], halo_hi: Optional[torch.Tensor]):
        lw, o, hp = self.w.layers[i], self.ops, self.hp
        di, nd = hp.d_inner, self.w.n_dbl
        lo = halo_lo.reshape(1, 3, di) if halo_lo is not None else None      # contiguous slices of the gather buffer
        hi = halo_hi.reshape(1, 3, di) if halo_hi is not None else None
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
        ws = self.ws
        if "h_end_c" in ws:
            h_end, sdl = ws["h_end_c"], ws["sdl_c"]
        else:
            h_end, sdl = ws["h_end"], ws["sdl"]
        self._scan(i, h_out=h_end, sum_delta=sdl, summary_only=True)
        return h_end, sdl

    def scan_summary_packed(self, i: int) -> torch.Tensor:
        """Summary pass into this rank's packed record ``[2*cmax*di*17]`` (one all-gather moves it)."""
        ws = self.ws
        h_end, sdl = self.scan_summary(i)
        if "h_end_c" in ws:
            ws["h_end"][:, :self.C].copy_(h_end)
            ws["sdl"][:, :self.C].copy_(sdl)
        return ws["pack"]

    def fold_packed(self, i: int, pack_all: torch.Tensor, g0: int, n_out: int) -> torch.Tensor:
        return self.ops.fold_states_packed(pack_all, self.w.layers[i]["A2"], self.world, self.cmax, self.di, g0, n_out,
                                           h_in=self.ws["h_in"])

    def fold(self, i: int, h_end: torch.Tensor, sdl: torch.Tensor, g0: int, n_out: int, h0=None, want_final=False,
             dir_mask: int = 3):
        return self.ops.fold_states(h_end.contiguous(), sdl.contiguous(), self.w.layers[i]["A2"], g0, n_out, h0=h0,
                                    want_final=want_final, dir_mask=dir_mask)

    def scan_seeded(self, i: int, h_in: torch.Tensor):
        self._scan(i, y=self.y, h_in=h_in.contiguous())

    def out_proj(self, i: int):
        lw, hp = self.w.layers[i], self.hp
        # residual add in the epilogue (MTN_EPI_RESADD without planes / row sums): same additions as Add -> Norm, bit-identical,
        # and the block output never makes its round trip through HBM (engine.LayerPlan._layer)
        self.ops.gemm(self.y, lw["w_out"], self.Lr, hp.d_model, 2 * hp.d_inner, out=self.res, epilogue=self._lib.EPI_RESADD,
                      epi_param=1)

    # ---- tail
    def head(self):
        hp, w, o = self.hp, self.w, self.ops
        o.add_rmsnorm(None, self.res, True, w.norm_f, self.P, xn=self.xn, beta=w.norm_f_b)
        o.gemm(self.xn, w.w_mask, self.Lr, hp.n_spk * hp.enc_dim, hp.d_model, out=self.sep_full[1:],
               epilogue=self._lib.EPI_MASK, epi_param=hp.enc_dim, aux=self.mix_w)

    def sep_last_row(self) -> torch.Tensor:
        self.ws["row"].copy_(self.sep_full[self.Lr])
        return self.ws["row"]

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
        est = self.ops.decoder(self.sep_full, self.w.w_dec, 1, T_x, Lx, hp.enc_dim, hp.n_spk, est=self.ws["est"],
                               frames=self.ws["frames"])
        return est[0]


# ------------------------------------------------------------------------------------------------ driver
class SequenceParallelSeparator:
    """``forward(mix [1, T]) -> est [1, T, n_spk]`` with time sharded over the ranks of ``group``.

    ``mix`` must be the same (replicated) tensor on every rank; the result is assembled on every rank.
    ``forward_local(mix_local, T)`` is the sharded form: rank r passes only the samples ``input_range(T)`` it owns and
    gets back the samples ``output_range(T)`` of the estimate -- no rank ever holds the whole recording (this is what an
    end-to-end pipeline with host buffers should call: every rank copies 1/world of the audio in and out).
    ``exchange``: ``"allgather"`` (default: the conv edges and ONE packed record of chunk summaries per layer) or
    ``"sendrecv"`` (the folded state is handed down the rank chain with point-to-point send/recv; same numbers).
    Collectives per forward (allgather mode): 2 per layer + 1 (decoder seam); ``collectives_per_forward`` reports it."""

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
        # The whole chunked forward -- collectives included (NCCL >= 2.9 is capturable) -- replays as one CUDA graph per
        # length.  On one rank that is also the low-latency plan for ONE short utterance: the scan's serial chain is
        # sub_chunks times shorter than in the batch plan (4 s @ 8 kHz, S: 9.0 ms -> 2.1 ms at 16 sub-chunks).
        # The point-to-point chain (exchange="sendrecv") stays eager when world > 1.
        cuda_be = backend is None or isinstance(backend, CudaSeqBackend)
        self.use_graph = use_graph and cuda_be and (self.comm.world == 1 or exchange == "allgather")
        self._graphs = LRUDict()   # one whole-forward graph (with its static input / output) per recording length
        self.graph_failed = None   # set to the error text if a multi-rank capture was refused (the driver then runs eagerly)

    def close(self):
        """Drop the captured graphs.  Call this (or delete the object) BEFORE ``dist.destroy_process_group()``: a live CUDA
        graph that holds captured NCCL collectives keeps the communicator busy and the teardown waits for it forever
        (observed on B200 / NCCL 2.28: tools/seqpar_check.py hung at exit until it released the driver first)."""
        self._graphs.clear()

    def __del__(self):
        try:
            self._graphs.clear()
        except Exception:
            pass

    # ---- geometry
    def plan(self, T: int) -> SeqPlan:
        return make_seq_plan(self.hp.frames(T), self.comm.world, self.sub_chunks)

    def input_range(self, T: int, rank: Optional[int] = None) -> Tuple[int, int]:
        """Samples ``[s0, s1)`` of the recording that rank ``rank`` encodes (its frames plus the 8-sample frame overlap)."""
        f0, f1 = self.plan(T).ranges[self.comm.rank if rank is None else rank]
        return 8 * f0, 8 * f1 + 8

    def output_range(self, T: int, rank: Optional[int] = None) -> Tuple[int, int]:
        """Samples ``[o0, o1)`` of the estimate that rank ``rank`` finalises (clipped to T: pad / trim of the reference)."""
        r = self.comm.rank if rank is None else rank
        f0, f1 = self.plan(T).ranges[r]
        hi = 8 * f1 + (8 if r == self.comm.world - 1 else 0)
        return min(8 * f0, T), (T if r == self.comm.world - 1 else min(hi, T))

    @property
    def collectives_per_forward(self) -> int:
        if self.comm.world == 1:
            return 0
        per_layer = 2 if self.exchange == "allgather" else 5      # edges + packed summaries | edges + 2 sends + 2 recvs
        return self.hp.n_mamba * per_layer + 1

    # ---- public API
    @torch.no_grad()
    def forward_local(self, mix_local: torch.Tensor, T: int) -> torch.Tensor:
        """``mix_local`` [1, s1 - s0] = this rank's ``input_range(T)`` of the recording -> ``[o1 - o0, n_spk]`` = this
        rank's ``output_range(T)`` of the estimate."""
        s0, s1 = self.input_range(T)
        if mix_local.dim() != 2 or mix_local.shape[0] != 1 or mix_local.shape[1] != s1 - s0:
            raise ValueError(f"rank {self.comm.rank} owns samples [{s0}, {s1}) of a {T}-sample recording: expected "
                             f"[1, {s1 - s0}], got {tuple(mix_local.shape)}")
        if not (self.use_graph and mix_local.is_cuda and self.graph_failed is None):
            return self._forward_local(mix_local, T)
        ent = self._graphs.get(T)
        if ent is None:
            static_in = mix_local.clone()
            out = self._forward_local(static_in, T)      # eager first: workspaces, kernel attributes, shape checks
            torch.cuda.current_stream().synchronize()
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    static_out = self._forward_local(static_in, T)
            except Exception as e:                       # pragma: no cover - depends on the NCCL / driver build
                if self.comm.world == 1:
                    raise
                self.graph_failed = f"{type(e).__name__}: {e}"
                torch.cuda.synchronize()
                return out
            ent = self._graphs[T] = (g, static_in, static_out)
        g, static_in, static_out = ent
        static_in.copy_(mix_local, non_blocking=True)
        g.replay()
        return static_out.clone()

    @torch.no_grad()
    def forward(self, mix: torch.Tensor) -> torch.Tensor:
        if mix.dim() != 2 or mix.shape[0] != 1:
            raise ValueError("sequence-parallel mode separates one recording: mix must be [1, T]")
        T = mix.shape[1]
        s0, s1 = self.input_range(T)
        piece = self.forward_local(mix[:, s0:s1].contiguous(), T)
        W = self.comm.world
        if W == 1:
            return piece.unsqueeze(0)
        ranges = [self.output_range(T, q) for q in range(W)]
        nmax = max(b - a for a, b in ranges)
        pad = piece.new_zeros((nmax, piece.shape[1]))
        pad[: piece.shape[0]] = piece
        allp = self.comm.all_gather(pad)                                     # result gather (replicated output only)
        return torch.cat([allp[q, : b - a] for q, (a, b) in enumerate(ranges)], dim=0).unsqueeze(0)

    __call__ = forward

    def _gather(self, t: torch.Tensor, name: str) -> torch.Tensor:
        buf = self.be.gather_buffer(name) if hasattr(self.be, "gather_buffer") else None
        return self.comm.all_gather(t, out=buf)

    def _forward_local(self, mix_local: torch.Tensor, T: int) -> torch.Tensor:
        hp, be, comm = self.hp, self.be, self.comm
        r, W = comm.rank, comm.world
        plan = self.plan(T)
        f0, f1 = plan.ranges[r]
        Lr, C, cmax = f1 - f0, plan.chunks[r], plan.cmax
        packed = self.exchange == "allgather" and hasattr(be, "scan_summary_packed")
        if hasattr(be, "set_plan"):
            be.set_plan(cmax, W)
        be.begin(mix_local, Lr, plan.Ls, C, plan.last_len[r])
        for i in range(be.n_layers):
            be.pre(i)
            edges = self._gather(be.xs_edges(), "edges_all")                 # [W, 2, 3, di]
            be.conv_xproj(i, edges[r - 1, 1] if r > 0 else None, edges[r + 1, 0] if r < W - 1 else None)
            if packed:
                pack_all = self._gather(be.scan_summary_packed(i), "pack_all")   # [W, 2*cmax*di*17]: ONE collective
                h_in = be.fold_packed(i, pack_all, r * cmax, C)
            elif self.exchange == "allgather":
                h_end, sdl = be.scan_summary(i)                              # [2, C, di, 16], [2, C, di]
                he = h_end.new_zeros((2, cmax) + tuple(h_end.shape[2:]))
                sd = sdl.new_zeros((2, cmax) + tuple(sdl.shape[2:]))
                he[:, :C], sd[:, :C] = h_end, sdl                            # padding chunks = identity operators
                rec = comm.all_gather(torch.cat([he.reshape(-1), sd.reshape(-1)]))   # [W, rec]: one collective
                n_h = he.numel()
                he_all = rec[:, :n_h].reshape((W, 2, cmax) + tuple(h_end.shape[2:])).transpose(0, 1)
                sd_all = rec[:, n_h:].reshape((W, 2, cmax) + tuple(sdl.shape[2:])).transpose(0, 1)
                h_in, _ = be.fold(i, he_all.reshape((2, W * cmax) + tuple(h_end.shape[2:])),
                                  sd_all.reshape((2, W * cmax) + tuple(sdl.shape[2:])), r * cmax, C)
            else:
                h_end, sdl = be.scan_summary(i)
                h_in = self._fold_chain(i, h_end, sdl, C)
            be.scan_seeded(i, h_in)
            be.out_proj(i)
        be.head()
        rows = self._gather(be.sep_last_row(), "rows_all")                   # [W, n_spk*N]
        be.set_sep_halo(rows[r - 1] if r > 0 else None)
        est_x = be.decode()                                                  # samples 8*(f0-1) .. 8*(f1+1)
        o0, o1 = self.output_range(T)
        n_keep = 8 * Lr + (8 if r == W - 1 else 0)                           # what this rank's frames finalise
        piece = est_x[8: 8 + min(n_keep, o1 - o0)]
        if piece.shape[0] < o1 - o0:                                         # T beyond the last frame: zero pad
            piece = torch.cat([piece, piece.new_zeros((o1 - o0 - piece.shape[0], piece.shape[1]))], dim=0)
        return piece

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
