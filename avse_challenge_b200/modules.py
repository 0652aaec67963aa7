"""``nn.Module`` drop-ins for the reference's Mamba-TasNet modules, backed by the sm_100a kernels.

Constructor signatures, parameter names/shapes (``state_dict`` keys) and call
conventions mirror the reference, so its checkpoints load with ``strict=True`` and
``Separation.compute_forward`` (``Mamba-TasNet/train_wsj0mix.py:86-111``) runs unchanged on them:

    Encoder(kernel_size, out_channels)      speechbrain dual_path.Encoder (yaml mambatasnet_S.yaml:131-133)
    MaskNet(enc_dim, bot_dim, ...)          modules/mamba_masknet.py:46-64
    Decoder(in_channels, out_channels, kernel_size, stride, bias)   dual_path.Decoder (yaml :151-156)
    MambaBlocksSequential / Block / Mamba   modules/mamba_blocks.py:108-123, modules/mamba/bimamba.py:410-412,40-61

Inference only (``torch.no_grad``); options the kernels do not implement raise at construction --
there is no silent fallback.  The fast path a user should call is ``MambaTasNetSeparator`` (one fused
plan, CUDA graph); the stand-alone module ``forward``s exist for drop-in compatibility and use the same
kernels with reference layouts ([B, N, L] in / out).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib, ops
from .engine import SeparatorEngine
from .hparams import HParams


def _unsupported(what):
    raise NotImplementedError(f"{what} is not implemented by the B200 kernels (no fallback path exists)")


class RMSNorm(nn.Module):
    """Parameter container for mamba-ssm's RMSNorm (``weight [D]``, no bias, eps 1e-5)."""

    def __init__(self, hidden_size, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden_size))


class ChannelwiseLayerNorm(nn.Module):
    """Parameter container for speechbrain's cLN (``gamma``/``beta`` ``[1,1,N]``, eps 1e-8)."""

    def __init__(self, channel_size):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(1, 1, channel_size))
        self.beta = nn.Parameter(torch.zeros(1, 1, channel_size))


class _SBConv1d(nn.Module):
    """speechbrain ``nnet.CNN.Conv1d`` (k=1, no bias): parameter lives at ``conv.weight [out, in, 1]``."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv1d(in_channels, out_channels, 1, bias=False)


class Mamba(nn.Module):
    """Mamba mixer parameters.  ``bimamba_type="v2"``: the bidirectional mixer of ``modules/mamba/bimamba.py:40-174``.
    ``bimamba_type="none"``: the unidirectional ``mamba_ssm.Mamba`` [3P] that ``modules/mamba_blocks.py:128`` selects for
    ``bidirectional=False`` -- same parameters without the ``*_b`` set; it also offers the reference's single-token
    decode interface (``step`` / ``allocate_inference_cache``, vendored copy at ``bimamba.py:320-380``)."""

    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", conv_bias=True, bias=False,
                 layer_idx=None, bimamba_type="v2", if_devide_out=True, init_layer_scale=None, **_ignored):
        super().__init__()
        if bimamba_type not in ("v2", "none"):
            _unsupported(f"bimamba_type={bimamba_type!r}")
        self.bimamba_type = bimamba_type
        if d_state != 16 or d_conv != 4:
            _unsupported(f"d_state={d_state}, d_conv={d_conv} (kernels are specialised for 16 / 4)")
        if bias or not conv_bias or not if_devide_out or init_layer_scale is not None:
            _unsupported("bias=True / conv_bias=False / if_devide_out=False / init_layer_scale")
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.layer_idx = layer_idx
        di, R = self.d_inner, self.dt_rank
        self.in_proj = nn.Linear(d_model, 2 * di, bias=False)
        self.conv1d = nn.Conv1d(di, di, d_conv, groups=di, padding=d_conv - 1, bias=True)
        self.x_proj = nn.Linear(di, R + 2 * d_state, bias=False)
        self.dt_proj = nn.Linear(R, di, bias=True)
        A_log = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32)).repeat(di, 1)
        self.A_log = nn.Parameter(A_log.clone())
        self.D = nn.Parameter(torch.ones(di))
        if bimamba_type == "v2":
            self.A_b_log = nn.Parameter(A_log.clone())
            self.conv1d_b = nn.Conv1d(di, di, d_conv, groups=di, padding=d_conv - 1, bias=True)
            self.x_proj_b = nn.Linear(di, R + 2 * d_state, bias=False)
            self.dt_proj_b = nn.Linear(R, di, bias=True)
            self.D_b = nn.Parameter(torch.ones(di))
        self.out_proj = nn.Linear(di, d_model, bias=False)
        self.__dict__["_packed"] = None

    # ------------------------------------------------------------------ single-token decode (unidirectional only)
    def allocate_inference_cache(self, batch_size, max_seqlen=None, dtype=None, **kwargs):
        """``(conv_state [B, di, d_conv], ssm_state [B, di, d_state])`` zeros, as ``bimamba.py:368-380``."""
        dev = self.out_proj.weight.device
        return (torch.zeros(batch_size, self.d_inner, self.d_conv, device=dev, dtype=dtype or torch.float32),
                torch.zeros(batch_size, self.d_inner, self.d_state, device=dev, dtype=dtype or torch.float32))

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.__dict__["_packed"] = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.__dict__["_packed"] = None
        return out

    @torch.no_grad()
    def step(self, hidden_states, conv_state, ssm_state):
        """``Mamba.step`` (``bimamba.py:320-372``): one token ``[B, 1, D]``; ``conv_state`` / ``ssm_state`` (fp32) are
        updated in place; returns ``(out [B, 1, D], conv_state, ssm_state)``.  Runs the same kernels as the sequence
        forward on a 1-frame chunk (in_proj, conv with the cached history, x_proj, scan seeded with the state, out_proj)."""
        if self.bimamba_type != "none":
            _unsupported("step() on the bidirectional mixer (the backward direction needs the future)")
        if hidden_states.dim() != 3 or hidden_states.shape[1] != 1:
            raise AssertionError("Only support decoding with 1 token at a time for now")   # bimamba.py:322
        if conv_state.dtype != torch.float32 or ssm_state.dtype != torch.float32:
            _unsupported("inference caches other than fp32")
        from .engine import pack_layer
        B, di, R, D = hidden_states.shape[0], self.d_inner, self.dt_rank, self.d_model
        dev = hidden_states.device
        hp = HParams("step", D, D, 1, d_state=self.d_state, expand=self.expand, d_conv=self.d_conv, bidirectional=False)
        lw = self.__dict__.get("_packed")
        if lw is None:
            sd = {"l.mixer." + k: v for k, v in self.state_dict().items()}
            sd["l.norm.weight"] = torch.ones(D, device=dev)
            lw = pack_layer(sd, "l.", hp, 2, dev)
            self.__dict__["_packed"] = lw
        nd = ops.n_dbl_for(R)
        xn = ops.split_planes(hidden_states.reshape(B, D).float().contiguous(), 2)
        xz = ops.gemm(xn, lw["w_in"], B, 2 * di, D, epilogue=_lib.EPI_INPROJ, epi_param=di)
        halo = conv_state[:, :, 1:].transpose(1, 2).contiguous()                    # the 3 inputs before this token
        u = ops.conv_silu(xz, lw["conv_w"], lw["conv_b"], B, 1, di, 2, halo_lo=halo, dir_mask=1)
        conv_state.copy_(torch.cat([conv_state[:, :, 1:], xz[:, :di].unsqueeze(-1)], dim=-1))   # bimamba.py:328-329
        dbl = torch.empty((B, 2 * nd), dtype=torch.float32, device=dev)
        ops.gemm(u, lw["w_x"], B, nd, di, out=dbl, groups=1, out_group_stride=nd)
        h = torch.zeros((2, B, di, 16), dtype=torch.float32, device=dev)
        h[0].copy_(ssm_state)
        y = ops.scan(u, dbl, xz, di, lw["w_dt"], lw["dt_bias"], lw["A2"], lw["D"], B, 1, di, R, dir_mask=1, h_in=h, h_out=h)
        ssm_state.copy_(h[0])
        out = ops.gemm(y, lw["w_out"], B, D, di)
        return out.view(B, 1, D).to(hidden_states.dtype), conv_state, ssm_state


class LayerNormParams(nn.LayerNorm):
    """``nn.LayerNorm`` parameter container (``rms_norm=False``: ``modules/mamba_blocks.py:36-41``)."""


class Block(nn.Module):
    """Add -> RMSNorm -> Mixer (``modules/mamba/bimamba.py:409-462``); parameters only."""

    def __init__(self, dim, mixer_cls, norm_cls=RMSNorm, fused_add_norm=False, residual_in_fp32=False):
        super().__init__()
        self.mixer = mixer_cls(dim)
        self.norm = norm_cls(dim)


def _stack_forward(self, x, keep_to=None, inference_params=None):
    """``MambaBlocksSequential.forward`` (``modules/mamba_blocks.py:186-212``): ``[B, L, D] -> [B, L, D]``.
    ``inference_params`` (unidirectional stacks): an object with ``seqlen_offset`` and ``key_value_memory_dict`` as the
    reference's mixers use it (``bimamba.py:186-190,382-404``): layer i's ``(conv_state [B, di, 4], ssm_state [B, di, 16])``
    are created on first use, read as the history of this call and updated in place -- prefill and ``step`` are the same
    chunk kernels here, so any number of tokens per call is accepted."""
    from .engine import MambaStack
    cache = self.__dict__.setdefault("_stack_cache", {})
    dev = self.norm_f.weight.device
    if "s" not in cache:
        D = self.norm_f.weight.shape[0]
        hp = HParams("stack", D, D, self.n_mamba, bidirectional=self.bidirectional, rms_norm=self.rms_norm)
        cache["s"] = MambaStack(hp, self.state_dict(), device=dev, mode="fp32")
    stack = cache["s"]
    B, L, D = x.shape
    states = None
    if inference_params is not None:
        if self.bidirectional:
            _unsupported("inference_params with bidirectional=True (the backward direction needs the future)")
        kv = inference_params.key_value_memory_dict
        di = stack.hp.d_inner
        # seqlen_offset == 0 is the reference's prefill: it starts from a zero state and OVERWRITES the caches
        # (bimamba.py:271-304), which is how callers reset a cache object for a new sequence
        fresh = getattr(inference_params, "seqlen_offset", 0) == 0
        from . import stream_fused
        if L <= stream_fused.MAX_FRAMES and x.is_cuda and stream_fused.eligible(stack.hp, stack.mode):
            # decode calls (a few tokens): ONE cluster-kernel launch for the whole stack.  The caches stay the reference's
            # tensors -- kv[i] = (conv_state [B, di, 4], ssm_state [B, di, 16]) -- but as views of two stacked buffers the
            # kernel indexes by layer; caches the caller already holds are adopted (copied in once, then replaced by views).
            if "f" not in cache:
                cache["f"] = stream_fused.FusedStack(stack)
            own = kv.get("_mtn_b200")
            if own is None or own["conv"].shape[1] != B or own["conv"].device != dev:
                conv = torch.zeros(self.n_mamba, B, 4, di, device=dev)
                ssm = torch.zeros(self.n_mamba, B, di, 16, device=dev)
                for i in range(self.n_mamba):
                    if i in kv:
                        conv[i].copy_(kv[i][0].transpose(1, 2))
                        ssm[i].copy_(kv[i][1])
                    kv[i] = (conv[i].transpose(1, 2), ssm[i])
                own = kv["_mtn_b200"] = {"conv": conv, "ssm": ssm}
            if fresh:
                own["conv"].zero_()
                own["ssm"].zero_()
            return cache["f"].run(x.float().contiguous(), own["conv"], own["ssm"]).to(x.dtype)
        states = []
        for i in range(self.n_mamba):
            if i not in kv:
                kv[i] = (torch.zeros(B, di, 4, device=dev), torch.zeros(B, di, 16, device=dev))
            cs, ss = kv[i]
            if fresh:
                cs.zero_()
                ss.zero_()
            conv4 = cs.transpose(1, 2).contiguous().float()
            h = torch.zeros(2, B, di, 16, device=dev)
            h[0].copy_(ss)
            states.append({"conv4": conv4, "halo": conv4[:, 1:].contiguous(), "h": h})
    out = stack.forward(x.float().contiguous(), states)
    if states is not None:
        for i, st in enumerate(states):
            cs, ss = inference_params.key_value_memory_dict[i]
            cs.copy_(st["conv4"].transpose(1, 2))
            ss.copy_(st["h"][0])
    return out.to(x.dtype)


class MambaBlocksSequential(nn.Module):
    """``modules/mamba_blocks.py:87-212`` (parameters + init); forward runs inside the fused engine."""

    def __init__(self, n_mamba, bidirectional=False, d_model=256, d_state=16, expand=2, d_conv=4, dt_rank="auto",
                 conv_bias=True, bias=False, fused_add_norm=True, rms_norm=False, norm_epsilon=1e-5,
                 initializer_cfg=None, residual_in_fp32=False):
        super().__init__()
        if norm_epsilon != 1e-5:
            _unsupported("norm_epsilon != 1e-5")
        # fused_add_norm only selects between two mathematically identical reference code paths
        # (mamba_blocks.py:195-210); residual_in_fp32: the residual stream here is always fp32.
        self.n_mamba, self.bidirectional, self.rms_norm = n_mamba, bidirectional, rms_norm
        btype = "v2" if bidirectional else "none"      # mamba_blocks.py:128: BiMamba if bidirectional else mamba_ssm.Mamba
        norm_cls = RMSNorm if rms_norm else (lambda d: LayerNormParams(d, eps=norm_epsilon))   # mamba_blocks.py:36-41
        mk = lambda i: Block(d_model, lambda d: Mamba(d, d_state=d_state, d_conv=d_conv, expand=expand, dt_rank=dt_rank,
                                                      conv_bias=conv_bias, bias=bias, layer_idx=i, bimamba_type=btype),
                             norm_cls=norm_cls)
        self.layers = nn.Sequential(*[mk(i) for i in range(n_mamba)])
        self.norm_f = RMSNorm(d_model, eps=norm_epsilon) if rms_norm else LayerNormParams(d_model, eps=norm_epsilon)
        with torch.no_grad():  # out_proj rescale, mamba_blocks.py:76-84
            for blk in self.layers:
                nn.init.kaiming_uniform_(blk.mixer.out_proj.weight, a=math.sqrt(5))
                blk.mixer.out_proj.weight /= math.sqrt(n_mamba)
                for dtp in ((blk.mixer.dt_proj, blk.mixer.dt_proj_b) if bidirectional else (blk.mixer.dt_proj,)):  # bimamba.py:101-118
                    R = dtp.weight.shape[1]
                    nn.init.uniform_(dtp.weight, -R ** -0.5, R ** -0.5)
                    dt = torch.exp(torch.rand(dtp.bias.shape[0]) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3)).clamp(min=1e-4)
                    dtp.bias.copy_(dt + torch.log(-torch.expm1(-dt)))

    forward = torch.no_grad()(_stack_forward)

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.__dict__["_stack_cache"] = {}

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.__dict__["_stack_cache"] = {}
        return out


class _EngineOwner(nn.Module):
    """Shared machinery: lazily build a ``SeparatorEngine`` from the current parameters."""

    def _invalidate(self):
        self.__dict__["_engine_cache"] = {}

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._invalidate()

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._invalidate()
        return out


class Encoder(nn.Module):
    """``relu(conv1d(mix, k=16, s=8))``: [B, T] -> [B, N, L] (view of the kernel's channel-last buffer)."""

    def __init__(self, kernel_size=16, out_channels=256, in_channels=1):
        super().__init__()
        if kernel_size != 16 or in_channels != 1:
            _unsupported("Encoder kernel_size != 16 or in_channels != 1")
        self.conv1d = nn.Conv1d(in_channels, out_channels, kernel_size, stride=kernel_size // 2, bias=False)
        self.in_channels = in_channels

    @torch.no_grad()
    def forward(self, x):
        N = self.conv1d.weight.shape[0]
        dev = x.device
        ones = torch.ones(N, device=dev)
        zeros = torch.zeros(N, device=dev)
        mix_w, _ = ops.encoder_cln(x.contiguous().float(), self.conv1d.weight.detach().reshape(N, 16).contiguous(),
                                   ones, zeros, 1)
        B, T = x.shape
        return mix_w.view(B, -1, N).transpose(1, 2)


class Decoder(nn.ConvTranspose1d):
    """``ConvTranspose1d(N -> 1, k=16, s=8, no bias)``: [B, N, L] -> [B, (L-1)*8+16]."""

    def __init__(self, in_channels, out_channels=1, kernel_size=16, stride=8, bias=False, **kw):
        if out_channels != 1 or kernel_size != 16 or stride != 8 or bias:
            _unsupported("Decoder with out_channels != 1 / kernel 16 / stride 8 / bias")
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, bias=False)

    @torch.no_grad()
    def forward(self, x):
        if x.dim() != 3:
            raise RuntimeError(f"Decoder expects [B, N, L], got {tuple(x.shape)}")
        B, N, L = x.shape
        sep = x.transpose(1, 2).contiguous().float().view(B * L, N)
        T_est = (L - 1) * 8 + 16
        est = ops.decoder(sep, self.weight.detach().reshape(N, 16).contiguous(), B, T_est, L, N, n_spk=1)
        # speechbrain's Decoder squeezes the channel dim only: [1, 1, T] -> [1, T] (same code as baseline/avse2/model.py:31-36:
        # squeeze(x).dim() == 1 -> squeeze(x, dim=1)), so compute_forward's `.unsqueeze(-1)` / `[:, :T_origin, :]` work at B = 1
        return est.view(B, T_est)


class MaskNet(_EngineOwner):
    """``modules/mamba_masknet.py:13-139``: [B, N, L] -> est_mask [n_spk, B, N, L] (ReLU mask)."""

    def __init__(self, enc_dim, bot_dim, n_spk=2, norm_type="gLN", causal=False, mask_nonlinear="relu", n_mamba=16,
                 bidirectional=True, d_model=256, d_state=16, expand=2, d_conv=4, fused_add_norm=False, rms_norm=True,
                 residual_in_fp32=False, mode="fp32"):
        super().__init__()
        if mask_nonlinear not in ("relu", "softmax"):
            raise ValueError("Unsupported mask non-linear function")      # modules/mamba_masknet.py:138
        if n_spk < 1:
            raise ValueError("n_spk must be >= 1")
        if bot_dim != d_model:
            _unsupported("bot_dim != d_model")
        self.n_spk, self.mask_nonlinear, self.mode = n_spk, mask_nonlinear, mode
        self.layer_norm = ChannelwiseLayerNorm(enc_dim)
        self.bottleneck_conv1x1 = _SBConv1d(enc_dim, bot_dim)
        self.mamba_net = MambaBlocksSequential(n_mamba=n_mamba, bidirectional=bidirectional, d_model=d_model,
                                               d_state=d_state, expand=expand, d_conv=d_conv,
                                               fused_add_norm=fused_add_norm, rms_norm=rms_norm,
                                               residual_in_fp32=residual_in_fp32, conv_bias=True, bias=False)
        self.mask_conv1x1 = _SBConv1d(bot_dim, n_spk * enc_dim)
        self.hp = HParams("custom", enc_dim, d_model, n_mamba, d_state=d_state, expand=expand, d_conv=d_conv, n_spk=n_spk,
                          bidirectional=bidirectional, mask_nonlinear=mask_nonlinear, rms_norm=rms_norm)
        self._invalidate()

    def engine(self, encoder_sd=None, decoder_sd=None, mode=None, use_graph=True) -> SeparatorEngine:
        mode = mode or self.mode
        key = (mode, use_graph, id(encoder_sd), id(decoder_sd))
        cache = self.__dict__.setdefault("_engine_cache", {})
        if key not in cache:
            dev = self.layer_norm.gamma.device
            N = self.hp.enc_dim
            enc = encoder_sd or {"conv1d.weight": torch.zeros(N, 1, 16)}
            dec = decoder_sd or {"weight": torch.zeros(N, 1, 16)}
            cache[key] = SeparatorEngine(self.hp, {"encoder": enc, "masknet": self.state_dict(), "decoder": dec},
                                         device=dev, mode=mode, use_graph=use_graph)
        return cache[key]

    @torch.no_grad()
    def forward(self, mixture_w):
        B, N, L = mixture_w.shape
        eng = self.engine(use_graph=False)
        hp, w, P = eng.hp, eng.w, eng.w.P
        x = mixture_w.transpose(1, 2).contiguous().float().view(B * L, N)
        ws = eng.workspace(B, (L - 1) * 8 + 16)
        ops.cln(x, w.gamma, w.beta, P, yn=ws.yn)
        ops.gemm(ws.yn, w.w_bot, ws.M, hp.d_model, N, out=ws.h)
        for i, lw in enumerate(w.layers):
            eng._layer(ws, lw, first=(i == 0))
        ops.add_rmsnorm(ws.h, ws.res, True, w.norm_f, P, xn=ws.xn, beta=w.norm_f_b)
        if self.mask_nonlinear == "softmax":                      # mamba_masknet.py:133-134 (dim 2 = the N channels)
            score = ops.softmax_mask(ops.gemm(ws.xn, w.w_mask, ws.M, hp.n_spk * N, hp.d_model), None, ws.M, N, hp.n_spk)
        else:
            score = ops.gemm(ws.xn, w.w_mask, ws.M, hp.n_spk * N, hp.d_model, epilogue=_lib.EPI_RELU)
        return score.view(B, L, hp.n_spk, N).permute(2, 0, 3, 1)  # mamba_masknet.py:126-131


class MambaTasNetSeparator(_EngineOwner):
    """Fused ``Encoder -> MaskNet -> mask * mix_w -> Decoder -> pad/trim`` == ``compute_forward``
    (``Mamba-TasNet/train_wsj0mix.py:86-111``); ``forward(mix [B, T]) -> est_source [B, T, n_spk]``."""

    def __init__(self, encoder: Encoder, masknet: MaskNet, decoder: Decoder, mode="fp32", use_graph=True):
        super().__init__()
        self.encoder, self.masknet, self.decoder = encoder, masknet, decoder
        self.mode, self.use_graph = mode, use_graph
        self._invalidate()

    @classmethod
    def from_hparams(cls, hp: HParams, mode="fp32", use_graph=True):
        enc = Encoder(hp.kernel_size, hp.enc_dim)
        mask = MaskNet(hp.enc_dim, hp.d_model, n_spk=hp.n_spk, n_mamba=hp.n_mamba, d_model=hp.d_model,
                       d_state=hp.d_state, expand=hp.expand, d_conv=hp.d_conv, mode=mode, bidirectional=hp.bidirectional,
                       mask_nonlinear=hp.mask_nonlinear, rms_norm=hp.rms_norm)
        dec = Decoder(hp.enc_dim, 1, hp.kernel_size, hp.stride, bias=False)
        return cls(enc, mask, dec, mode=mode, use_graph=use_graph)

    def load_reference_state_dicts(self, sds: dict, strict=True):
        """``sds = {"encoder": ..., "masknet": ..., "decoder": ...}`` as saved by the reference's checkpointer
        (``inference.ipynb`` cell 1)."""
        self.encoder.load_state_dict(sds["encoder"], strict=strict)
        self.masknet.load_state_dict(sds["masknet"], strict=strict)
        self.decoder.load_state_dict(sds["decoder"], strict=strict)
        self._invalidate()
        return self

    def engine(self) -> SeparatorEngine:
        cache = self.__dict__.setdefault("_engine_cache", {})
        if "e" not in cache:
            dev = self.masknet.layer_norm.gamma.device
            sds = {"encoder": self.encoder.state_dict(), "masknet": self.masknet.state_dict(),
                   "decoder": self.decoder.state_dict()}
            cache["e"] = SeparatorEngine(self.masknet.hp, sds, device=dev, mode=self.mode, use_graph=self.use_graph)
        return cache["e"]

    @torch.no_grad()
    def forward(self, mix):
        return self.engine().forward(mix)

    def chunked(self, sub_chunks: int = 16):
        """The chunked-scan plan over this module's weights (``parallel.SequenceParallelSeparator``): ``[1, T] ->
        [1, T, n_spk]``.  One utterance at minimum latency on one GPU (the whole forward is one CUDA graph; the scan's
        serial chain is ``sub_chunks`` times shorter than in the batch plan), or one long recording sharded over the
        ranks of the default process group.  Bidirectional stacks only."""
        from .parallel import SequenceParallelSeparator
        cache = self.__dict__.setdefault("_engine_cache", {})
        key = ("chunked", sub_chunks)
        if key not in cache:
            dev = self.masknet.layer_norm.gamma.device
            sds = {"encoder": self.encoder.state_dict(), "masknet": self.masknet.state_dict(),
                   "decoder": self.decoder.state_dict()}
            cache[key] = SequenceParallelSeparator(self.masknet.hp, sds, device=dev, mode=self.mode,
                                                   sub_chunks=sub_chunks, use_graph=self.use_graph)
        return cache[key]


# ---------------------------------------------------------------------------------------------------- DPMamba
class Dual_Computation_Block(nn.Module):
    """Parameter container of speechbrain's ``Dual_Computation_Block`` (``norm="ln"``, no linear layers)."""

    def __init__(self, intra_mdl, inter_mdl, out_channels, norm="ln", skip_around_intra=True,
                 linear_layer_after_inter_intra=True):
        super().__init__()
        if norm != "ln" or linear_layer_after_inter_intra:
            _unsupported("Dual_Computation_Block with norm != 'ln' or linear_layer_after_inter_intra=True")
        self.intra_mdl, self.inter_mdl = intra_mdl, inter_mdl
        self.skip_around_intra = skip_around_intra
        self.intra_norm = nn.GroupNorm(1, out_channels, eps=1e-8)
        self.inter_norm = nn.GroupNorm(1, out_channels, eps=1e-8)


class Dual_Path_Model(_EngineOwner):
    """speechbrain ``lobes.models.dual_path.Dual_Path_Model`` as the dpmamba recipes instantiate it
    (``hparams/WSJ0Mix/dpmamba_L.yaml:164-174``): same constructor, same state_dict keys; ``forward`` is
    [B, N, L] -> [n_spk, B, N, L] (``modules/dual_path.py:56-150``).  ``intra_model`` / ``inter_model`` must be this
    package's one-layer-per-stack bidirectional ``MambaBlocksSequential``."""

    def __init__(self, in_channels, out_channels, intra_model, inter_model, num_layers=1, norm="ln", K=200, num_spks=2,
                 skip_around_intra=True, linear_layer_after_inter_intra=True, use_global_pos_enc=False,
                 max_length=20000, mode="fp32"):
        super().__init__()
        import copy
        if use_global_pos_enc:
            _unsupported("use_global_pos_enc=True")
        for mdl in (intra_model, inter_model):
            if not isinstance(mdl, MambaBlocksSequential) or not mdl.bidirectional:
                _unsupported("intra / inter models other than a bidirectional MambaBlocksSequential")
        if intra_model.n_mamba != inter_model.n_mamba:
            _unsupported("intra / inter stacks of different depth")
        if num_spks < 1 or K % 2:
            _unsupported("num_spks < 1 or odd K")
        self.K, self.num_spks, self.num_layers, self.mode = K, num_spks, num_layers, mode
        self.norm = nn.GroupNorm(1, in_channels, eps=1e-8)
        self.conv1d = nn.Conv1d(in_channels, out_channels, 1, bias=False)
        self.dual_mdl = nn.ModuleList([
            copy.deepcopy(Dual_Computation_Block(intra_model, inter_model, out_channels, norm,
                                                 skip_around_intra=skip_around_intra,
                                                 linear_layer_after_inter_intra=linear_layer_after_inter_intra))
            for _ in range(num_layers)])
        self.conv2d = nn.Conv2d(out_channels, out_channels * num_spks, kernel_size=1)
        self.end_conv1x1 = nn.Conv1d(out_channels, in_channels, 1, bias=False)
        self.prelu = nn.PReLU()
        self.output = nn.Sequential(nn.Conv1d(out_channels, out_channels, 1), nn.Tanh())
        self.output_gate = nn.Sequential(nn.Conv1d(out_channels, out_channels, 1), nn.Sigmoid())
        from .hparams import DPHParams
        self.hp = DPHParams("custom_dp", in_channels, out_channels, num_layers, skip_around_intra, chunk_size=K,
                            n_mamba_dp=2 * intra_model.n_mamba, n_spk=num_spks)
        self._invalidate()

    def engine(self, encoder_sd=None, decoder_sd=None, mode=None, use_graph=True):
        from .dpmamba import DPSeparatorEngine
        mode = mode or self.mode
        key = (mode, use_graph, id(encoder_sd), id(decoder_sd))
        cache = self.__dict__.setdefault("_engine_cache", {})
        if key not in cache:
            N = self.hp.enc_dim
            enc = encoder_sd or {"conv1d.weight": torch.zeros(N, 1, 16)}
            dec = decoder_sd or {"weight": torch.zeros(N, 1, 16)}
            cache[key] = DPSeparatorEngine(self.hp, {"encoder": enc, "masknet": self.state_dict(), "decoder": dec},
                                           device=self.conv1d.weight.device, mode=mode, use_graph=use_graph)
        return cache[key]

    @torch.no_grad()
    def forward(self, x):
        B, N, L = x.shape
        eng = self.engine(use_graph=False)
        ws = eng.workspace(B, (L - 1) * 8 + 16)
        ws.mix_w.copy_(x.transpose(1, 2).reshape(B * L, N))
        mask = eng._run(ws, mask_only=True)
        return mask.view(B, L, self.num_spks, N).permute(2, 0, 3, 1).clone()


class Dual_Path_Model_Skip(Dual_Path_Model):
    """``modules/dual_path.py:17-150`` (vendored subclass): every ``skip_n_block`` dual blocks the running tensor is
    averaged with the segmented input (``:114-116``).  ``skip_n_block = 0`` -- what every shipped recipe sets
    (``dpmamba_L.yaml:116``) -- makes it identical to ``Dual_Path_Model``."""

    def __init__(self, *args, skip_n_block=0, **kw):
        if skip_n_block < 0:
            raise ValueError("skip_n_block must be >= 0")
        super().__init__(*args, **kw)
        self.skip_n_block = skip_n_block
        from dataclasses import replace
        self.hp = replace(self.hp, skip_n_block=skip_n_block)


class DPMambaSeparator(_EngineOwner):
    """Fused ``Encoder -> Dual_Path_Model -> mask * mix_w -> Decoder`` (``train_wsj0mix.py:86-111`` with the dpmamba
    recipes): ``forward(mix [B, T]) -> est_source [B, T, n_spk]``."""

    def __init__(self, encoder: Encoder, masknet: Dual_Path_Model, decoder: Decoder, mode="fp32", use_graph=True):
        super().__init__()
        self.encoder, self.masknet, self.decoder = encoder, masknet, decoder
        self.mode, self.use_graph = mode, use_graph
        self._invalidate()

    @classmethod
    def from_hparams(cls, hp, mode="fp32", use_graph=True):
        mk = lambda: MambaBlocksSequential(hp.n_mamba_dp // 2, bidirectional=True, d_model=hp.d_model, d_state=hp.d_state,
                                           expand=hp.expand, d_conv=hp.d_conv, fused_add_norm=False, rms_norm=True)
        kw = dict(num_layers=hp.n_dp, norm="ln", K=hp.chunk_size, num_spks=hp.n_spk, skip_around_intra=hp.skip_around_intra,
                  linear_layer_after_inter_intra=False, mode=mode)
        if hp.skip_n_block:   # modules/dual_path.py:17-150 (no shipped recipe sets it)
            mask = Dual_Path_Model_Skip(hp.enc_dim, hp.d_model, mk(), mk(), skip_n_block=hp.skip_n_block, **kw)
        else:
            mask = Dual_Path_Model(hp.enc_dim, hp.d_model, mk(), mk(), **kw)
        return cls(Encoder(hp.kernel_size, hp.enc_dim), mask, Decoder(hp.enc_dim, 1, hp.kernel_size, hp.stride, bias=False),
                   mode=mode, use_graph=use_graph)

    def load_reference_state_dicts(self, sds: dict, strict=True):
        self.encoder.load_state_dict(sds["encoder"], strict=strict)
        self.masknet.load_state_dict(sds["masknet"], strict=strict)
        self.decoder.load_state_dict(sds["decoder"], strict=strict)
        self._invalidate()
        self.masknet._invalidate()
        return self

    def engine(self):
        from .dpmamba import DPSeparatorEngine
        cache = self.__dict__.setdefault("_engine_cache", {})
        if "e" not in cache:
            sds = {"encoder": self.encoder.state_dict(), "masknet": self.masknet.state_dict(),
                   "decoder": self.decoder.state_dict()}
            cache["e"] = DPSeparatorEngine(self.masknet.hp, sds, device=self.masknet.conv1d.weight.device, mode=self.mode,
                                           use_graph=self.use_graph)
        return cache["e"]

    @torch.no_grad()
    def forward(self, mix):
        return self.engine().forward(mix)
