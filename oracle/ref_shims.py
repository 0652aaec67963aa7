"""TEST INFRASTRUCTURE ONLY -- import the REAL reference modules under leaf shims.

The reference's hot path (``/root/reference/Mamba-TasNet/modules``) hard-imports
third-party packages that are not installed here (``causal_conv1d``,
``causal_conv1d_cuda``, ``selective_scan_cuda``, ``mamba_ssm``, ``speechbrain``;
see ``modules/mamba/selective_scan_interface.py:14-16``,
``modules/mamba_masknet.py:3-8``, ``modules/mamba_blocks.py:12-17``).  This module
registers minimal stand-ins for exactly those *leaves* in ``sys.modules`` so that
the reference's own ``MaskNet`` / ``MambaBlocksSequential`` / ``Block`` / bi-``Mamba``
/ ``MambaInnerFnNoOutProj`` / ``selective_scan_ref`` code runs unmodified on CPU.

Leaf semantics (each is the documented behaviour of the pinned third-party
version, ``Mamba-TasNet/requirement.txt:7-11``):

* ``causal_conv1d_cuda.causal_conv1d_fwd(x, w, b, seq_idx, silu)``
  = depthwise causal conv (left zero pad ``width-1``) + bias + optional SiLU;
  same arithmetic the reference uses in its non-CUDA branch
  (``modules/mamba/bimamba.py:279``).
* ``selective_scan_cuda.fwd(...)`` -> routed to the reference's own
  ``selective_scan_ref`` (``modules/mamba/selective_scan_interface.py:91-157``).
* ``mamba_ssm.ops.triton.layernorm.RMSNorm``: ``x * rsqrt(mean(x^2) + eps) * w``
  in fp32 (mamba-ssm 1.1.3.post1 ``rms_norm_ref``).
* ``speechbrain.lobes.models.conv_tasnet.ChannelwiseLayerNorm``: per-token
  mean/var over channels (biased), ``gamma*(y-mean)/sqrt(var+1e-8)+beta``,
  params ``gamma``/``beta`` of shape ``[1,1,N]`` (speechbrain 1.0.0).
* ``speechbrain.nnet.CNN.Conv1d``: channel-last wrapper around ``nn.Conv1d``
  stored as child ``conv`` (state_dict key ``conv.weight``).

This file can only work where ``/root/reference`` exists (the build container).
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("MTN_REFERENCE_ROOT", "/root/reference")
_MT_ROOT = os.path.join(REFERENCE_ROOT, "Mamba-TasNet")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(_MT_ROOT, "modules"))


# --------------------------------------------------------------------------- leaves
def _no_autocast():
    """The third-party kernels compute in fp32 whatever dtype they are handed and round ONCE on output; their stand-ins
    must not be re-cast by a surrounding autocast region (the bf16 golden run of oracle/make_golden.py)."""
    return torch.autocast("cpu", enabled=False)


def _causal_conv1d_fwd(x, weight, bias, seq_idx, silu):
    # x [B, D, L]; weight [D, W].  causal-conv1d 1.1.3.post1: fp32 accumulation, output in x.dtype.
    d, w = weight.shape
    with _no_autocast():
        y = F.conv1d(x.float(), weight.float()[:, None, :], None if bias is None else bias.float(), padding=w - 1,
                     groups=d)[..., : x.shape[-1]]
        return (F.silu(y) if silu else y).to(x.dtype)


class _RMSNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-5, device=None, dtype=None):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden_size, device=device, dtype=dtype))
        self.register_parameter("bias", None)

    def forward(self, x):
        with _no_autocast():
            xf = x.float()
            y = xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + self.eps) * self.weight.float()
            return y.to(x.dtype)


class _ChannelwiseLayerNorm(nn.Module):
    def __init__(self, channel_size):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(1, 1, channel_size))
        self.beta = nn.Parameter(torch.zeros(1, 1, channel_size))

    def forward(self, y):
        mean = torch.mean(y, dim=2, keepdim=True)
        var = torch.var(y, dim=2, keepdim=True, unbiased=False)
        return self.gamma * (y - mean) / torch.pow(var + 1e-8, 0.5) + self.beta


class _SBConv1d(nn.Module):
    """speechbrain.nnet.CNN.Conv1d restricted to what MaskNet uses (k=1, no bias)."""

    def __init__(self, out_channels, kernel_size, in_channels=None, bias=True, **kw):
        super().__init__()
        assert kernel_size == 1
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, bias=bias)

    def forward(self, x):  # [B, L, C] -> [B, L, C_out]
        return self.conv(x.transpose(1, -1)).transpose(1, -1)


class _DualComputationBlock(nn.Module):
    """speechbrain 1.0.0 ``lobes.models.dual_path.Dual_Computation_Block`` [3P, restated from the published source]
    restricted to ``linear_layer_after_inter_intra=False`` (all dpmamba recipes, ``dpmamba_L.yaml:173``)."""

    def __init__(self, intra_mdl, inter_mdl, out_channels, norm="ln", skip_around_intra=True,
                 linear_layer_after_inter_intra=True):
        super().__init__()
        assert not linear_layer_after_inter_intra and norm == "ln"
        self.intra_mdl, self.inter_mdl = intra_mdl, inter_mdl
        self.skip_around_intra = skip_around_intra
        self.intra_norm = nn.GroupNorm(1, out_channels, eps=1e-8)      # select_norm("ln", out_channels, 4)
        self.inter_norm = nn.GroupNorm(1, out_channels, eps=1e-8)

    def forward(self, x):
        B, N, K, S = x.shape
        intra = x.permute(0, 3, 2, 1).contiguous().view(B * S, K, N)
        intra = self.intra_mdl(intra)
        intra = intra.view(B, S, K, N).permute(0, 3, 2, 1).contiguous()
        intra = self.intra_norm(intra)
        if self.skip_around_intra:
            intra = intra + x
        inter = intra.permute(0, 2, 3, 1).contiguous().view(B * K, S, N)
        inter = self.inter_mdl(inter)
        inter = inter.view(B, K, S, N).permute(0, 3, 1, 2).contiguous()
        inter = self.inter_norm(inter)
        return inter + intra


class _DualPathModel(nn.Module):
    """speechbrain 1.0.0 ``lobes.models.dual_path.Dual_Path_Model`` [3P, restated from the published source]: members,
    ``_padding`` / ``_Segmentation`` / ``_over_add`` and ``forward``.  The forward is additionally pinned by the vendored
    subclass ``Mamba-TasNet/modules/dual_path.py:56-150`` (same statements), which the golden generator runs on top of
    this class."""

    def __init__(self, in_channels, out_channels, intra_model, inter_model, num_layers=1, norm="ln", K=200, num_spks=2,
                 skip_around_intra=True, linear_layer_after_inter_intra=True, use_global_pos_enc=False, max_length=20000):
        super().__init__()
        import copy
        assert not use_global_pos_enc
        self.K, self.num_spks, self.num_layers = K, num_spks, num_layers
        self.use_global_pos_enc = use_global_pos_enc
        self.norm = nn.GroupNorm(1, in_channels, eps=1e-8)             # select_norm(norm, in_channels, 3)
        self.conv1d = nn.Conv1d(in_channels, out_channels, 1, bias=False)
        self.dual_mdl = nn.ModuleList([
            copy.deepcopy(_DualComputationBlock(intra_model, inter_model, out_channels, norm,
                                                skip_around_intra=skip_around_intra,
                                                linear_layer_after_inter_intra=linear_layer_after_inter_intra))
            for _ in range(num_layers)])
        self.conv2d = nn.Conv2d(out_channels, out_channels * num_spks, kernel_size=1)
        self.end_conv1x1 = nn.Conv1d(out_channels, in_channels, 1, bias=False)
        self.prelu = nn.PReLU()
        self.activation = nn.ReLU()
        self.output = nn.Sequential(nn.Conv1d(out_channels, out_channels, 1), nn.Tanh())
        self.output_gate = nn.Sequential(nn.Conv1d(out_channels, out_channels, 1), nn.Sigmoid())

    def forward(self, x):
        x = self.norm(x)
        x = self.conv1d(x)
        x, gap = self._Segmentation(x, self.K)
        for i in range(self.num_layers):
            x = self.dual_mdl[i](x)
        x = self.prelu(x)
        x = self.conv2d(x)
        B, _, K, S = x.shape
        x = x.view(B * self.num_spks, -1, K, S)
        x = self._over_add(x, gap)
        x = self.output(x) * self.output_gate(x)
        x = self.end_conv1x1(x)
        _, N, L = x.shape
        x = x.view(B, self.num_spks, N, L)
        x = self.activation(x)
        return x.transpose(0, 1)

    def _padding(self, input, K):
        B, N, L = input.shape
        P = K // 2
        gap = K - (P + L % K) % K
        if gap > 0:
            input = torch.cat([input, torch.zeros(B, N, gap).type(input.type())], dim=2)
        _pad = torch.zeros(B, N, P).type(input.type())
        return torch.cat([_pad, input, _pad], dim=2), gap

    def _Segmentation(self, input, K):
        B, N, L = input.shape
        P = K // 2
        input, gap = self._padding(input, K)
        input1 = input[:, :, :-P].contiguous().view(B, N, -1, K)
        input2 = input[:, :, P:].contiguous().view(B, N, -1, K)
        input = torch.cat([input1, input2], dim=3).view(B, N, -1, K).transpose(2, 3)
        return input.contiguous(), gap

    def _over_add(self, input, gap):
        B, N, K, S = input.shape
        P = K // 2
        input = input.transpose(2, 3).contiguous().view(B, N, -1, K * 2)
        input1 = input[:, :, :, :K].contiguous().view(B, N, -1)[:, :, P:]
        input2 = input[:, :, :, K:].contiguous().view(B, N, -1)[:, :, :-P]
        input = input1 + input2
        if gap > 0:
            input = input[:, :, :-gap]
        return input


_loaded = None


def load_reference():
    """Return a namespace with the reference's own classes/functions (CPU-runnable)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    warnings.filterwarnings("ignore", category=FutureWarning)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("causal_conv1d", causal_conv1d_fn=None, causal_conv1d_update=None)
    mod("causal_conv1d_cuda", causal_conv1d_fwd=_causal_conv1d_fwd)
    ssc = mod("selective_scan_cuda")
    mod("mamba_ssm", Mamba=None)
    mod("mamba_ssm.ops")
    mod("mamba_ssm.ops.triton")
    mod("mamba_ssm.ops.triton.layernorm", RMSNorm=_RMSNorm, layer_norm_fn=None, rms_norm_fn=None)
    sb = mod("speechbrain")
    sb.nnet = mod("speechbrain.nnet")
    sb.nnet.CNN = mod("speechbrain.nnet.CNN", Conv1d=_SBConv1d)
    mod("speechbrain.lobes")
    mod("speechbrain.lobes.models")
    mod("speechbrain.lobes.models.conv_tasnet", ChannelwiseLayerNorm=_ChannelwiseLayerNorm)
    mod("speechbrain.lobes.models.dual_path", Dual_Path_Model=_DualPathModel)

    sys.dont_write_bytecode = True  # reference tree is read-only
    if _MT_ROOT not in sys.path:
        sys.path.insert(0, _MT_ROOT)
    from modules.mamba import selective_scan_interface as ssi  # noqa: E402

    def _fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus):
        with _no_autocast():   # selective_scan_ref floats u / delta / B / C and returns out.to(u.dtype), like the kernel;
            # z is the one input it does not float (ssi.py:155 would take silu in bf16), while the kernel gates in fp32
            out_z, last = ssi.selective_scan_ref(u, delta, A, B, C, D, None if z is None else z.float(), delta_bias,
                                                 delta_softplus, True)
        # `x` = the kernel's per-chunk running states [B, D, n_chunks, 2*Ns]; SelectiveScanFn.forward reads the final
        # state from x[:, :, -1, 1::2] (selective_scan_interface.py:46)
        x = torch.zeros(last.shape[0], last.shape[1], 1, 2 * last.shape[2], dtype=last.dtype)
        x[:, :, 0, 1::2] = last
        return out_z, x, out_z

    ssc.fwd = _fwd
    from modules.mamba.bimamba import Mamba as BiMamba, Block  # noqa: E402

    class UniMamba(BiMamba):
        """Stand-in for ``mamba_ssm.Mamba`` (mamba-ssm 1.1.3.post1 ``modules/mamba_simple.py``, not vendored; used by
        ``modules/mamba_blocks.py:128`` when ``bidirectional=False``).  The vendored ``bimamba.Mamba`` is a modified
        copy of that class: its non-fused branch (``bimamba.py:271-310``) and ``step`` (``:320-372``) are the
        unidirectional mixer verbatim, using only the forward-direction parameters.  So: construct the vendored class,
        drop the ``*_b`` parameters (state_dict keys then equal ``mamba_ssm.Mamba``'s), and force the non-fused branch."""

        def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", conv_bias=True, bias=False,
                     use_fast_path=True, layer_idx=None, device=None, dtype=None, **kw):
            super().__init__(d_model, d_state=d_state, d_conv=d_conv, expand=expand, dt_rank=dt_rank,
                             conv_bias=conv_bias, bias=bias, layer_idx=layer_idx, device=device, dtype=dtype,
                             bimamba_type="v2")
            for name in ("A_b_log", "conv1d_b", "x_proj_b", "dt_proj_b", "D_b"):
                delattr(self, name)
            self.use_fast_path = False   # bimamba.py:203 -> falls through to the unidirectional branch at :271

    sys.modules["mamba_ssm"].Mamba = UniMamba
    from modules.mamba_masknet import MaskNet  # noqa: E402
    from modules.mamba_blocks import MambaBlocksSequential  # noqa: E402

    import contextlib
    import io
    from modules.dual_path import Dual_Path_Model_Skip  # noqa: E402  (vendored subclass: its forward is the pinned one)

    class InferenceParams:
        """mamba_ssm.utils.generation.InferenceParams restricted to the two fields the vendored code reads
        (``bimamba.py:186-190,375-404``)."""

        def __init__(self, max_seqlen=0, max_batch_size=0):
            self.max_seqlen, self.max_batch_size = max_seqlen, max_batch_size
            self.seqlen_offset = 0
            self.key_value_memory_dict = {}

    class RefEncoder(nn.Module):
        """speechbrain dual_path.Encoder == baseline/avse2/model.py:14-24 (in-repo twin)."""

        def __init__(self, kernel_size=16, out_channels=256, in_channels=1):
            super().__init__()
            self.conv1d = nn.Conv1d(in_channels, out_channels, kernel_size,
                                    stride=kernel_size // 2, groups=1, bias=False)

        def forward(self, x):
            return F.relu(self.conv1d(x.unsqueeze(1)))

    class RefDecoder(nn.ConvTranspose1d):
        """speechbrain dual_path.Decoder == baseline/avse2/model.py:27-37."""

        def forward(self, x):
            x = super().forward(x)
            return torch.squeeze(x, dim=1) if torch.squeeze(x).dim() == 1 else torch.squeeze(x)

    def compute_forward(encoder, masknet, decoder, mix, num_spks=2):
        """Line-for-line behaviour of Separation.compute_forward
        (Mamba-TasNet/train_wsj0mix.py:86-111), calling the modules passed in."""
        mix_w = encoder(mix)
        est_mask = masknet(mix_w)
        mix_w = torch.stack([mix_w] * num_spks)
        sep_h = mix_w * est_mask
        est_source = torch.cat([decoder(sep_h[i]).unsqueeze(-1) for i in range(num_spks)], dim=-1)
        if est_source.dim() == 2:  # batch 1 got squeezed by the decoder
            est_source = est_source.unsqueeze(0)
        t_origin, t_est = mix.size(1), est_source.size(1)
        if t_origin > t_est:
            est_source = F.pad(est_source, (0, 0, 0, t_origin - t_est))
        else:
            est_source = est_source[:, :t_origin, :]
        return est_source

    _loaded = types.SimpleNamespace(
        MaskNet=MaskNet, MambaBlocksSequential=MambaBlocksSequential, BiMamba=BiMamba, Block=Block,
        selective_scan_ref=ssi.selective_scan_ref, Encoder=RefEncoder, Decoder=RefDecoder,
        compute_forward=compute_forward, ssi=ssi, UniMamba=UniMamba, InferenceParams=InferenceParams,
        Dual_Path_Model=_DualPathModel, Dual_Path_Model_Skip=Dual_Path_Model_Skip,
    )
    return _loaded


def build_reference_model(hp, seed=1234, bidirectional=True, mask_nonlinear="relu", rms_norm=True):
    """Construct reference Encoder/MaskNet/Decoder with the reference's own init under a seed.

    ``hp`` needs: enc_dim, d_model, n_mamba, kernel_size."""
    ref = load_reference()
    torch.manual_seed(seed)
    enc = ref.Encoder(kernel_size=hp["kernel_size"], out_channels=hp["enc_dim"])
    mask = ref.MaskNet(enc_dim=hp["enc_dim"], bot_dim=hp["d_model"], n_spk=2, n_mamba=hp["n_mamba"],
                       mask_nonlinear=mask_nonlinear,
                       bidirectional=bidirectional, d_model=hp["d_model"], d_state=16, expand=2, d_conv=4,
                       fused_add_norm=False, rms_norm=rms_norm, residual_in_fp32=False)
    dec = ref.Decoder(in_channels=hp["enc_dim"], out_channels=1, kernel_size=hp["kernel_size"],
                      stride=hp["kernel_size"] // 2, bias=False)
    return enc.eval(), mask.eval(), dec.eval()


def build_reference_dp_model(hp, seed=1234):
    """DPMamba as ``hparams/WSJ0Mix/dpmamba_*.yaml:135-181`` builds it: Encoder, ``Dual_Path_Model`` with one
    ``MambaBlocksSequential(n_mamba_dp // 2)`` each as intra and inter model, Decoder.  The mask network is instantiated
    through the vendored ``Dual_Path_Model_Skip`` (``modules/dual_path.py``, ``skip_n_block=0``) so that the forward that
    runs is the one in the reference tree.  ``hp``: dict of a ``DPHParams``."""
    import contextlib
    import io
    ref = load_reference()
    torch.manual_seed(seed)
    enc = ref.Encoder(kernel_size=hp["kernel_size"], out_channels=hp["enc_dim"])
    mk = lambda: ref.MambaBlocksSequential(n_mamba=hp["n_mamba_dp"] // 2, bidirectional=True, d_model=hp["d_model"],
                                           d_state=16, expand=2, d_conv=4, fused_add_norm=False, rms_norm=True,
                                           residual_in_fp32=False)
    with contextlib.redirect_stdout(io.StringIO()):       # the vendored ctor prints skip_n_block
        mask = ref.Dual_Path_Model_Skip(in_channels=hp["enc_dim"], out_channels=hp["d_model"], intra_model=mk(),
                                        inter_model=mk(), num_layers=hp["n_dp"], norm="ln", K=hp["chunk_size"], num_spks=2,
                                        skip_around_intra=hp["skip_around_intra"], skip_n_block=hp.get("skip_n_block", 0),
                                        linear_layer_after_inter_intra=False)
    dec = ref.Decoder(in_channels=hp["enc_dim"], out_channels=1, kernel_size=hp["kernel_size"],
                      stride=hp["kernel_size"] // 2, bias=False)
    return enc.eval(), mask.eval(), dec.eval()
