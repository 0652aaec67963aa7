"""Causal (unidirectional) Mamba-TasNet as a streaming separator.

With ``bidirectional=False`` (``modules/mamba_blocks.py:128``) every stage of the separator is causal: the encoder
frame l sees samples ``[8l, 8l+16)``, cLN / RMSNorm / the 1x1 convs are per frame, the mixer's conv looks 3 frames back
and its scan carries a state.  The reference streams such a stack through ``inference_params`` caches -- a prefill call
(``modules/mamba/bimamba.py:271-304``) followed by ``Mamba.step`` one frame at a time (``:320-372``), each layer keeping
``conv_state [B, di, 4]`` and ``ssm_state [B, di, 16]``.  ``StreamingSeparator`` is that mechanism for the whole
waveform-to-waveform path and for chunks of any number of frames: the same kernels as the one-shot forward, seeded with

    * the last 8 input samples (the encoder's 16-sample window overlaps the previous chunk by half),
    * per layer: the last 3 conv inputs (``halo_lo`` of ``mtn_conv_silu_dir_fwd``) and the SSM state
      (``h_in`` / ``h_out`` of ``mtn_scan_fwd``, updated in place),
    * the second half of the last decoder frame (``tail`` of ``mtn_decoder_stream_fwd``).

Feeding a signal in chunks gives the same samples as separating it in one call; the algorithmic latency is one encoder
window (16 samples = 2 ms at 8 kHz).  One CUDA graph per chunk size replays the ~100 launches of a chunk.
"""
from __future__ import annotations

import torch

from . import _lib, stream_fused
from ._cache import LRUDict
from .engine import SeparatorEngine


class StreamingSeparator:
    """``push(chunk [B, 8*F]) -> est [B, 8*F', n_spk]`` with F' = F (F - 1 for the very first chunk, whose first frame
    needs 16 samples); ``flush()`` returns the last 8 samples (the tail of the final frame)."""

    def __init__(self, engine: SeparatorEngine, batch: int, use_graph: bool = True, fused: bool | None = None,
                 channels_per_cta: int | None = None):
        """``fused``: chunks of <= 32 frames go through the one-launch cluster kernel (``mtn_stream_push_fwd``) instead of
        the batch plan's ~100 launches.  None = whenever that kernel implements the configuration; True = require it."""
        if engine.hp.bidirectional:
            raise NotImplementedError("streaming needs a causal stack: construct the model with bidirectional=False")
        self.eng, self.batch, self.use_graph = engine, batch, use_graph
        self.channels_per_cta = channels_per_cta    # fused push: 32 / 64 / 128 d_inner channels per CTA; None = by batch size
        hp, dev = engine.hp, engine.device
        can_fuse = stream_fused.eligible(hp, engine.mode) and not engine.fuse_norm
        if fused and not can_fuse:
            raise _lib.MtnError("fused=True: the one-launch push does not implement this configuration / plan")
        self._fused = stream_fused.FusedPush(engine) if (can_fuse if fused is None else fused) else None
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        # one tensor per kind of cache (the fused kernel indexes it by layer); the batch plan sees per-layer views
        self._halo = z(hp.n_mamba, batch, 3, hp.d_inner)
        self._h = z(hp.n_mamba, 2, batch, hp.d_inner, 16)
        self.state = {
            "layers": [{"halo": self._halo[i], "h": self._h[i]} for i in range(hp.n_mamba)],
            "ola_tail": z(batch, hp.n_spk, 8),
            "est": None,
        }
        self.in_tail = z(batch, 8)
        self.started = False
        self.samples_in = 0
        self.samples_out = 0
        # per chunk length: output buffer + (graph, workspace it was captured against).  LRU-bounded; a graph is also
        # dropped when the engine has meanwhile evicted its workspace (checked at replay)
        self._est = LRUDict(on_evict=lambda L, buf: self._drop_graphs(L))
        self._graphs = LRUDict()

    def _drop_graphs(self, L):
        for key in [k for k, (_, ws) in self._graphs.items() if ws.L == L]:
            self._graphs.pop(key, None)

    def reset(self):
        for st in self.state["layers"]:
            st["halo"].zero_()
            st["h"].zero_()
        self.state["ola_tail"].zero_()
        self.in_tail.zero_()
        self.started = False
        self.samples_in = self.samples_out = 0

    def _run(self, ws, L):
        est = self._est.get(L)
        if est is None:
            est = self._est[L] = torch.empty((self.batch, 8 * L, self.eng.hp.n_spk), dtype=torch.float32,
                                             device=self.eng.device)
        self.state["est"] = est
        return self.eng._run(ws, stream_state=self.state)   # either plan (fuse_norm folds Add -> RMSNorm into the GEMMs)

    @torch.no_grad()
    def push(self, chunk: torch.Tensor) -> torch.Tensor:
        if chunk.dim() != 2 or chunk.shape[0] != self.batch or chunk.dtype != torch.float32 or not chunk.is_cuda:
            raise _lib.MtnError(f"push expects a CUDA fp32 tensor [batch={self.batch}, 8*F]")
        n = chunk.shape[1]
        if n % 8 != 0 or n == 0 or (not self.started and n < 16):
            raise _lib.MtnError(f"chunk of {n} samples: need a positive multiple of the hop (8), and >= 16 for the first chunk")
        T = n if not self.started else n + 8          # samples the encoder sees: [carried 8 | chunk]
        L = self.eng.hp.frames(T)
        if self._fused is not None and L <= stream_fused.MAX_FRAMES:
            # one cluster-kernel launch: it reads the chunk and the carried samples in place and updates every cache
            if chunk.stride(1) != 1:
                chunk = chunk.contiguous()
            with torch.cuda.device(self.eng.device):   # launches go to the engine's device, whatever the caller has current
                est = self._fused.run(chunk, self.in_tail, not self.started, self._halo, self._h, self.state["ola_tail"],
                                      dsl=self.channels_per_cta)
            self.started = True
            self.samples_in += n
            self.samples_out += 8 * L
            return est
        ws = self.eng.workspace(self.batch, T)
        if self.started:
            ws.mix[:, :8].copy_(self.in_tail)
            ws.mix[:, 8:T].copy_(chunk)
        else:
            ws.mix[:, :T].copy_(chunk)
        self.in_tail.copy_(chunk[:, n - 8:])
        key = (T,)
        if self.use_graph:
            g, g_ws = self._graphs.get(key, (None, None))
            if g is None or g_ws is not ws:   # never captured, or the engine evicted (and re-made) this shape's workspace
                # capture without running eagerly first: a warm-up run would advance the caches twice
                snap = self._snapshot()
                self._run(ws, L)
                torch.cuda.current_stream().synchronize()
                self._restore(snap)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run(ws, L)
                self._graphs[key] = (g, ws)
            g.replay()
            est = self._est[L]
        else:
            est = self._run(ws, L)
        self.started = True
        self.samples_in += n
        self.samples_out += 8 * L
        return est.clone()

    def flush(self) -> torch.Tensor:
        """The 8 samples still held in the overlap-add tail: ``[B, 8, n_spk]``.  Ends the stream (state is reset)."""
        out = self.state["ola_tail"].transpose(1, 2).clone()
        self.reset()
        return out

    def _snapshot(self):
        return ([(st["halo"].clone(), st["h"].clone()) for st in self.state["layers"]], self.state["ola_tail"].clone())

    def _restore(self, snap):
        layers, tail = snap
        for st, (halo, h) in zip(self.state["layers"], layers):
            st["halo"].copy_(halo)
            st["h"].copy_(h)
        self.state["ola_tail"].copy_(tail)

    @torch.no_grad()
    def separate(self, mix: torch.Tensor, chunk_samples: int) -> torch.Tensor:
        """Convenience: stream ``mix [B, T]`` through in chunks and return ``[B, T, n_spk]`` (zero-padded / trimmed to T
        like ``train_wsj0mix.py:104-109``).  T is cut to a multiple of 8 first (a trailing partial hop makes no frame)."""
        B, T = mix.shape
        self.reset()
        usable = T // 8 * 8
        outs, pos = [], 0
        while pos < usable:
            n = min(chunk_samples, usable - pos)
            if not self.started and n < 16:
                break
            outs.append(self.push(mix[:, pos:pos + n].contiguous()))
            pos += n
        outs.append(self.flush())
        est = torch.cat(outs, dim=1)
        if est.shape[1] < T:
            est = torch.nn.functional.pad(est, (0, 0, 0, T - est.shape[1]))
        return est[:, :T]
