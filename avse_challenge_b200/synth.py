"""Synthetic two-speaker mixtures (no dataset is reachable: WSJ0 is licensed, no network).

Recipe from SURVEY.md section 8(d): each source is a sum of 8 harmonics of an f0 random walk
in 90-250 Hz, amplitude-modulated by a 3-6 Hz raised-cosine "syllable" envelope with random
pauses, plus -30 dB white noise, scaled to RMS 0.05; ``mix = s1 + s2``.
Seed 1234 is the reference recipe's seed (``Mamba-TasNet/hparams/WSJ0Mix/mambatasnet_L.yaml:9-10``).
"""
from __future__ import annotations

import math

import torch


def synth_sources(batch: int, T: int, sample_rate: int = 8000, n_spk: int = 2, seed: int = 1234,
                  noise_second_source: bool = False) -> torch.Tensor:
    """Return sources ``[batch, T, n_spk]`` (fp32, CPU)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(T, dtype=torch.float64) / sample_rate
    out = torch.empty(batch, T, n_spk, dtype=torch.float32)
    hop = max(1, sample_rate // 100)  # f0 control rate 100 Hz
    n_ctl = T // hop + 2
    for b in range(batch):
        for s in range(n_spk):
            if noise_second_source and s == 1:
                x = torch.randn(T, generator=g, dtype=torch.float64)
            else:
                f0_0 = 90.0 + 160.0 * torch.rand(1, generator=g).item()
                walk = torch.cumsum(torch.randn(n_ctl, generator=g, dtype=torch.float64) * 1.5, 0)
                f0_ctl = (f0_0 + walk).clamp(90.0, 250.0)
                f0 = torch.nn.functional.interpolate(f0_ctl[None, None], size=n_ctl * hop,
                                                     mode="linear", align_corners=False)[0, 0, :T]
                phase = 2 * math.pi * torch.cumsum(f0, 0) / sample_rate
                x = torch.zeros(T, dtype=torch.float64)
                for h in range(1, 9):
                    amp = 1.0 / h * (0.5 + torch.rand(1, generator=g).item())
                    x = x + amp * torch.sin(h * phase + 2 * math.pi * torch.rand(1, generator=g).item())
                rate = 3.0 + 3.0 * torch.rand(1, generator=g).item()
                env = 0.5 * (1 - torch.cos(2 * math.pi * rate * t + 2 * math.pi * torch.rand(1, generator=g).item()))
                # random pauses: zero ~20 % of syllables
                syl = torch.floor(rate * t).long()
                keep = (torch.rand(int(syl.max().item()) + 2, generator=g) > 0.2).double()
                x = x * env * keep[syl]
                x = x + 10 ** (-30 / 20) * x.abs().mean().clamp(min=1e-6) * torch.randn(T, generator=g, dtype=torch.float64)
            x = x * (0.05 / x.pow(2).mean().sqrt().clamp(min=1e-9))
            out[b, :, s] = x.float()
    return out


def synth_mixture(batch: int, T: int, sample_rate: int = 8000, seed: int = 1234,
                  noise_second_source: bool = False):
    """Return ``(mix [batch, T], sources [batch, T, 2])``."""
    src = synth_sources(batch, T, sample_rate, 2, seed, noise_second_source)
    return src.sum(dim=-1), src


def si_snr(est: torch.Tensor, ref: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """Scale-invariant SNR in dB over the time axis (dim 1); inputs ``[B, T, C]``.

    Same definition as the reference's in-repo ``cal_si_snr``
    (``baseline/avse2/utils/dnn.py:15-57``): zero-mean, project, 10*log10, EPS 1e-8.
    """
    est = est.double() - est.double().mean(dim=1, keepdim=True)
    ref = ref.double() - ref.double().mean(dim=1, keepdim=True)
    dot = (est * ref).sum(dim=1, keepdim=True)
    energy = ref.pow(2).sum(dim=1, keepdim=True) + eps
    proj = dot * ref / energy
    noise = est - proj
    ratio = proj.pow(2).sum(dim=1) / (noise.pow(2).sum(dim=1) + eps)
    return 10 * torch.log10(ratio + eps)


def pit_si_snr(est: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """Best-permutation (2 speakers) mean SI-SNR per utterance, ``[B]``."""
    a = si_snr(est, src).mean(dim=-1)
    b = si_snr(est.flip(-1), src).mean(dim=-1)
    return torch.maximum(a, b)
