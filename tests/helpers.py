"""Shared test helpers (CPU side)."""
import numpy as np
import torch


def load_golden_forward(path):
    z = np.load(path)
    sds = {"encoder": {}, "masknet": {}, "decoder": {}}
    for k in z.files:
        if "/" in k and k.split("/", 1)[0] in sds:
            grp, name = k.split("/", 1)
            sds[grp][name] = torch.from_numpy(z[k])
    taps = {k.split("/", 1)[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith("tap/")}
    out = {k: torch.from_numpy(z[k]) for k in ("mix", "src", "est", "mix_w", "est_mask")}
    return sds, out, taps


def rel_max(a, b):
    """max |a-b| / rms(b) -- the north-star waveform metric."""
    return ((a.double() - b.double()).abs().max() / b.double().pow(2).mean().sqrt().clamp(min=1e-30)).item()


def rel_mixed(a, b):
    """max |a-b| / (|b| + rms(b)): element-wise relative error with an rms floor.  Used for per-op checks on
    heavy-tailed random data, where max-abs/rms is dominated by the rounding of a few very large elements."""
    b = b.double()
    return ((a.double() - b).abs() / (b.abs() + b.pow(2).mean().sqrt().clamp(min=1e-30))).max().item()


def hp_from_sds(sds):
    from avse_challenge_b200.hparams import HParams
    m = sds["masknet"]
    N = m["layer_norm.gamma"].shape[-1]
    D = m["mamba_net.norm_f.weight"].shape[0]
    n = 1 + max(int(k.split(".")[2]) for k in m if k.startswith("mamba_net.layers."))
    return HParams("golden", N, D, n, rms_norm="mamba_net.norm_f.bias" not in m)
