"""The four shipped Mamba-TasNet configurations and state_dict initialisers.

Shapes follow the reference yaml object graphs
(``Mamba-TasNet/hparams/WSJ0Mix/mambatasnet_{XS,S,M,L}.yaml:108-128``); the reference
needs HyperPyYAML to read those, this package hard-codes the resolved values
(SURVEY.md section 0 table).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict, replace

import torch


@dataclass(frozen=True)
class HParams:
    name: str
    enc_dim: int          # N_encoder_out
    d_model: int          # out_channels == bottleneck == Mamba d_model
    n_mamba: int
    kernel_size: int = 16
    d_state: int = 16
    expand: int = 2
    d_conv: int = 4
    n_spk: int = 2
    sample_rate: int = 8000
    mask_nonlinear: str = "relu"  # or "softmax" (modules/mamba_masknet.py:133-138)
    rms_norm: bool = True         # False: nn.LayerNorm blocks and norm_f (modules/mamba_blocks.py:36-41,167-169)
    bidirectional: bool = True   # False: causal stack of unidirectional mixers (mamba_blocks.py:128, `mamba_ssm.Mamba`)

    @property
    def stride(self) -> int:
        return self.kernel_size // 2

    @property
    def d_inner(self) -> int:
        return self.expand * self.d_model

    @property
    def dt_rank(self) -> int:
        return math.ceil(self.d_model / 16)

    def frames(self, T: int) -> int:
        return (T - self.kernel_size) // self.stride + 1

    def as_dict(self):
        return asdict(self)

    def causal(self) -> "HParams":
        """Same sizes with ``bidirectional=False`` (the streaming-capable variant)."""
        return replace(self, name=self.name + "_causal", bidirectional=False)


CONFIGS = {
    "XS": HParams("XS", 128, 128, 16),
    "S": HParams("S", 256, 256, 16),
    "M": HParams("M", 256, 256, 32),
    "L": HParams("L", 512, 512, 32),
    # not a shipped config: small shapes for unit tests / golden fixtures
    "tiny": HParams("tiny", 64, 64, 2),
}


def init_state_dicts(hp: HParams, seed: int = 1234, trained_like: bool = True):
    """Synthesise ``{encoder, masknet, decoder}`` state_dicts with the reference's key names,
    shapes and init distributions:

    * ``A_log = log(1..16)``, ``D = 1`` (``modules/mamba/bimamba.py:123-134``)
    * ``dt_proj.weight ~ U(+-dt_rank^-0.5)``, ``dt_proj.bias = softplus^-1(U_log[1e-3, 0.1])``
      (``bimamba.py:101-120``)
    * ``out_proj.weight`` kaiming-uniform / sqrt(n_layer) (``modules/mamba_blocks.py:76-84``)
    * everything else torch's default ``nn.Linear`` / ``nn.Conv1d`` init.

    ``trained_like=True`` additionally perturbs ``A_log``/``D``/norm weights so tests do not
    silently rely on the special structure of the untrained init (A = -(n+1), D = 1, g = 1).
    """
    g = torch.Generator().manual_seed(seed)
    N, D, di, R, Ns, W = hp.enc_dim, hp.d_model, hp.d_inner, hp.dt_rank, hp.d_state, hp.d_conv

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def linear_w(out_f, in_f):  # kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in))
        return uni((out_f, in_f), 1.0 / math.sqrt(in_f))

    enc = {"conv1d.weight": uni((N, 1, hp.kernel_size), 1.0 / math.sqrt(hp.kernel_size))}
    dec = {"weight": uni((N, 1, hp.kernel_size), 1.0 / math.sqrt(hp.kernel_size))}
    m = {}
    m["layer_norm.gamma"] = torch.ones(1, 1, N)
    m["layer_norm.beta"] = torch.zeros(1, 1, N)
    m["bottleneck_conv1x1.conv.weight"] = linear_w(D, N).unsqueeze(-1)
    for i in range(hp.n_mamba):
        p = f"mamba_net.layers.{i}."
        A_log = torch.log(torch.arange(1, Ns + 1, dtype=torch.float32)).repeat(di, 1)
        m[p + "mixer.A_log"] = A_log.clone()
        m[p + "mixer.D"] = torch.ones(di)
        if hp.bidirectional:
            m[p + "mixer.A_b_log"] = A_log.clone()
            m[p + "mixer.D_b"] = torch.ones(di)
        m[p + "mixer.in_proj.weight"] = linear_w(2 * di, D)
        for sfx in (("", "_b") if hp.bidirectional else ("",)):
            m[p + f"mixer.conv1d{sfx}.weight"] = uni((di, 1, W), 1.0 / math.sqrt(W))
            m[p + f"mixer.conv1d{sfx}.bias"] = uni((di,), 1.0 / math.sqrt(W))
            m[p + f"mixer.x_proj{sfx}.weight"] = linear_w(R + 2 * Ns, di)
            m[p + f"mixer.dt_proj{sfx}.weight"] = uni((di, R), R ** -0.5)
            dt = torch.exp(torch.rand(di, generator=g) * (math.log(0.1) - math.log(1e-3))
                           + math.log(1e-3)).clamp(min=1e-4)
            m[p + f"mixer.dt_proj{sfx}.bias"] = dt + torch.log(-torch.expm1(-dt))
        m[p + "mixer.out_proj.weight"] = linear_w(D, di) / math.sqrt(hp.n_mamba)
        m[p + "norm.weight"] = torch.ones(D)
        if not hp.rms_norm:
            m[p + "norm.bias"] = torch.zeros(D)
    m["mamba_net.norm_f.weight"] = torch.ones(D)
    if not hp.rms_norm:
        m["mamba_net.norm_f.bias"] = torch.zeros(D)
    m["mask_conv1x1.conv.weight"] = linear_w(hp.n_spk * N, D).unsqueeze(-1)
    if trained_like:
        for k in list(m.keys()):
            if k.endswith("A_log") or k.endswith("A_b_log"):
                m[k] = m[k] + 0.3 * torch.randn(m[k].shape, generator=g)
            elif k.endswith(".D") or k.endswith(".D_b"):
                m[k] = m[k] + 0.2 * torch.randn(m[k].shape, generator=g)
            elif k.endswith("norm.weight") or k.endswith("norm_f.weight") or k.endswith("gamma"):
                m[k] = m[k] + 0.1 * torch.randn(m[k].shape, generator=g)
            elif k.endswith("beta") or k.endswith("norm.bias") or k.endswith("norm_f.bias"):
                m[k] = m[k] + 0.05 * torch.randn(m[k].shape, generator=g)
    return {"encoder": enc, "masknet": m, "decoder": dec}


# ------------------------------------------------------------------------------------------------------- DPMamba
@dataclass(frozen=True)
class DPHParams:
    """DPMamba recipes (``Mamba-TasNet/hparams/WSJ0Mix/dpmamba_{XS,S,M,L}.yaml:108-123,164-174``): speechbrain
    ``Dual_Path_Model`` with ``MambaBlocksSequential(n_mamba_dp // 2)`` as intra and as inter model."""
    name: str
    enc_dim: int                 # N_encoder_out
    d_model: int                 # out_channels
    n_dp: int                    # dual-path blocks
    skip_around_intra: bool
    chunk_size: int = 250        # K
    n_mamba_dp: int = 2
    kernel_size: int = 16
    d_state: int = 16
    expand: int = 2
    d_conv: int = 4
    n_spk: int = 2
    sample_rate: int = 8000
    skip_n_block: int = 0        # Dual_Path_Model_Skip (modules/dual_path.py:114-116); 0 in every shipped recipe

    @property
    def stride(self) -> int:
        return self.kernel_size // 2

    def frames(self, T: int) -> int:
        return (T - self.kernel_size) // self.stride + 1

    @property
    def stack(self) -> HParams:
        """Hyper-parameters of one intra / inter stack."""
        return HParams(self.name + "_stack", self.d_model, self.d_model, self.n_mamba_dp // 2, kernel_size=self.kernel_size,
                       d_state=self.d_state, expand=self.expand, d_conv=self.d_conv, n_spk=self.n_spk,
                       sample_rate=self.sample_rate)

    def as_dict(self):
        return asdict(self)


DP_CONFIGS = {
    "XS": DPHParams("dp_XS", 128, 128, 8, False),
    "S": DPHParams("dp_S", 256, 256, 8, False),
    "M": DPHParams("dp_M", 256, 256, 16, True),
    "L": DPHParams("dp_L", 512, 512, 16, True),
    # not shipped: small shapes for unit tests / golden fixtures (K = 10 -> many ragged chunks at short lengths)
    "tiny": DPHParams("dp_tiny", 64, 64, 2, True, chunk_size=10),
}


def init_dp_state_dicts(hp: DPHParams, seed: int = 1234, trained_like: bool = True):
    """``{encoder, masknet, decoder}`` state_dicts of a DPMamba model with speechbrain ``Dual_Path_Model`` key names
    (``norm``, ``conv1d``, ``dual_mdl.<i>.{intra_mdl,inter_mdl,intra_norm,inter_norm}``, ``conv2d``, ``end_conv1x1``,
    ``prelu``, ``output.0``, ``output_gate.0``); the stacks are initialised like ``init_state_dicts``."""
    g = torch.Generator().manual_seed(seed)
    N, D = hp.enc_dim, hp.d_model

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    enc = {"conv1d.weight": uni((N, 1, hp.kernel_size), 1.0 / math.sqrt(hp.kernel_size))}
    dec = {"weight": uni((N, 1, hp.kernel_size), 1.0 / math.sqrt(hp.kernel_size))}
    m = {}
    pert = (lambda t, s: t + s * torch.randn(t.shape, generator=g)) if trained_like else (lambda t, s: t)
    m["norm.weight"], m["norm.bias"] = pert(torch.ones(N), 0.1), pert(torch.zeros(N), 0.05)
    m["conv1d.weight"] = uni((D, N, 1), 1.0 / math.sqrt(N))
    for i in range(hp.n_dp):
        for which in ("intra", "inter"):
            sub = init_state_dicts(hp.stack, seed=seed + 17 * i + (0 if which == "intra" else 7),
                                   trained_like=trained_like)["masknet"]
            for k, v in sub.items():
                if k.startswith("mamba_net."):
                    m[f"dual_mdl.{i}.{which}_mdl.{k[len('mamba_net.'):]}"] = v
            m[f"dual_mdl.{i}.{which}_norm.weight"] = pert(torch.ones(D), 0.1)
            m[f"dual_mdl.{i}.{which}_norm.bias"] = pert(torch.zeros(D), 0.05)
    m["conv2d.weight"] = uni((D * hp.n_spk, D, 1, 1), 1.0 / math.sqrt(D))
    m["conv2d.bias"] = uni((D * hp.n_spk,), 1.0 / math.sqrt(D))
    m["end_conv1x1.weight"] = uni((N, D, 1), 1.0 / math.sqrt(D))
    m["prelu.weight"] = torch.full((1,), 0.25)
    for name in ("output", "output_gate"):
        m[f"{name}.0.weight"] = uni((D, D, 1), 1.0 / math.sqrt(D))
        m[f"{name}.0.bias"] = uni((D,), 1.0 / math.sqrt(D))
    return {"encoder": enc, "masknet": m, "decoder": dec}
