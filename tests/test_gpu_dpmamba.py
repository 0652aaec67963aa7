"""GPU (-m gpu): DPMamba -- the dual-path glue kernels against the CPU oracle and the whole
separator against golden vectors minted from the reference's vendored Dual_Path_Model_Skip.forward
(oracle/make_golden.py).  Gates: max|est - ref| <= 1e-3 * rms(ref), |dSI-SNR| <= 0.01 dB."""
import os
from dataclasses import replace

import numpy as np
import pytest
import torch

from avse_challenge_b200 import DP_CONFIGS, init_dp_state_dicts, synth_mixture, pit_si_snr
from avse_challenge_b200 import _lib, ops, modules
from avse_challenge_b200.dpmamba import DPSeparatorEngine
from oracle import restate
from tests.helpers import load_golden_forward, rel_max

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _gate(est, ref, src):
    return rel_max(est, ref), (pit_si_snr(est, src) - pit_si_snr(ref, src)).abs().max().item()


@pytest.mark.parametrize("B,rows,C", [(2, 40, 64), (3, 8500, 256), (1, 3999, 512), (2, 5, 4)])
def test_group_norm_stats_and_apply(B, rows, C):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, rows, C, generator=g) * 3 + 1.5            # non-zero mean: exercises the variance formula
    w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = restate.group_norm1(x, w, b)
    xd = x.to(DEV)
    part = ops.gn_stats(xd, B, rows, C)
    out = torch.empty_like(xd)
    planes = torch.empty(2, B * rows, C, dtype=torch.bfloat16, device=DEV)
    ops.gn_apply(xd, part, w.to(DEV), b.to(DEV), B, 1, rows, C, out_a=out, planes=planes)
    assert rel_max(out.cpu(), ref) <= 1e-5
    assert rel_max(planes.float().sum(0).view(B, rows, C).cpu(), ref) <= 1e-4
    part2 = ops.gn_stats(xd, B, rows, C)
    assert torch.equal(part, part2)                               # deterministic reduction


def test_group_norm_apply_transposed_and_skip():
    g = torch.Generator().manual_seed(2)
    B, S, K, C = 2, 6, 10, 64
    x_t = torch.randn(B, K, S, C, generator=g)                    # rows (b, k, s), as the inter stack writes them
    skip = torch.randn(B, S, K, C, generator=g)
    w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = restate.group_norm1(x_t.transpose(1, 2), w, b) + skip   # rows (b, s, k)
    xd = x_t.contiguous().to(DEV)
    part = ops.gn_stats(xd, B, S * K, C)
    out_a, out_a2, out_t = (torch.empty(B, S, K, C, device=DEV), torch.empty(B, S, K, C, device=DEV),
                            torch.empty(B, K, S, C, device=DEV))
    ops.gn_apply(xd, part, w.to(DEV), b.to(DEV), B, S, K, C, skip=skip.to(DEV), out_a=out_a, out_a2=out_a2, out_t=out_t,
                 x_transposed=True)
    assert rel_max(out_a.cpu(), ref) <= 1e-5
    assert torch.equal(out_a, out_a2) and torch.equal(out_t, out_a.transpose(1, 2).contiguous())


@pytest.mark.parametrize("L,K", [(3999, 250), (7, 10), (10, 10), (249, 250), (375, 250)])
def test_segmentation_and_overlap_add(L, K):
    g = torch.Generator().manual_seed(3)
    B, C = 2, 64
    x = torch.randn(B, L, C, generator=g)
    seg_ref, gap = restate.dp_segment(x, K)
    S = ops.dp_num_chunks(L, K)
    assert S == seg_ref.shape[1]
    X, X2 = torch.empty(B, S, K, C, device=DEV), torch.empty(B, S, K, C, device=DEV)
    ops.dp_segment(x.to(DEV), B, L, C, K, S, X, X2)
    assert torch.equal(X.cpu(), seg_ref) and torch.equal(X, X2)
    y = torch.randn(B, S, K, C, generator=g)                      # arbitrary chunk contents (not a segmentation)
    a = torch.tensor([0.25])
    ref = restate.dp_over_add(torch.where(y >= 0, y, a * y), gap)
    planes = torch.empty(2, B * L, C, dtype=torch.bfloat16, device=DEV)
    ops.dp_overadd_prelu(y.to(DEV), a.to(DEV), planes, B, L, C, K, S)
    assert rel_max(planes.float().sum(0).view(B, L, C).cpu(), ref) <= 1e-4
    with pytest.raises(_lib.MtnError):
        ops.dp_segment(x.to(DEV), B, L, C, K, S + 2, X)


def test_bias_and_gate_planes():
    g = torch.Generator().manual_seed(4)
    rows, D, spk = 301, 64, 2
    x = torch.randn(rows, spk * D, generator=g)
    bias = torch.randn(spk * D, generator=g)
    planes = torch.empty(2, rows, spk * D, dtype=torch.bfloat16, device=DEV)
    ops.bias_planes(x.to(DEV), bias.to(DEV), 2.0, planes, rows, spk * D)
    assert rel_max(planes.float().sum(0).cpu(), x + 2 * bias) <= 1e-4
    og = torch.randn(rows, spk, 2 * D, generator=g) * 3
    bo, bg = torch.randn(D, generator=g), torch.randn(D, generator=g)
    ref = torch.tanh(og[..., :D] + bo) * torch.sigmoid(og[..., D:] + bg)
    gp = torch.empty(2, rows, spk * D, dtype=torch.bfloat16, device=DEV)
    ops.gate_planes(og.to(DEV), bo.to(DEV), bg.to(DEV), gp, rows, spk, D)
    assert (gp.float().sum(0).view(rows, spk, D).cpu() - ref).abs().max().item() <= 2e-5


@pytest.mark.parametrize("tag", ["dp_tiny_skip", "dp_tiny_noskip"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_dpmamba_end_to_end_matches_reference_golden(golden_dir, tag, use_graph):
    path = os.path.join(golden_dir, f"forward_{tag}.npz")
    sds, g, _ = load_golden_forward(path)
    z = np.load(path)
    hp = replace(DP_CONFIGS["tiny"], skip_around_intra=bool(z["skip_around_intra"]), chunk_size=int(z["chunk_size"]),
                 n_dp=int(z["n_dp"]))
    sep = modules.DPMambaSeparator.from_hparams(hp, mode="fp32", use_graph=use_graph)
    sep.load_reference_state_dicts(sds, strict=True).to(DEV)
    est = sep(g["mix"].to(DEV)).cpu()
    est2 = sep(g["mix"].to(DEV)).cpu()
    assert torch.equal(est, est2)
    err, d = _gate(est, g["est"], g["src"])
    print(f"{tag}: max-abs/rms {err:.3e} dSI-SNR {d:.2e}")
    assert err <= 1e-3 and d <= 0.01, (err, d)


def test_dual_path_model_skip_block_blend_matches_reference_golden(golden_dir):
    """Dual_Path_Model_Skip with skip_n_block = 1 (vendored modules/dual_path.py:114-116): golden from the reference."""
    path = os.path.join(golden_dir, "forward_dp_tiny_blockskip.npz")
    sds, g, _ = load_golden_forward(path)
    z = np.load(path)
    hp = replace(DP_CONFIGS["tiny"], n_dp=int(z["n_dp"]), skip_n_block=int(z["skip_n_block"]))
    est = DPSeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False)(g["mix"].to(DEV)).cpu()
    err, d = _gate(est, g["est"], g["src"])
    assert err <= 1e-3 and d <= 0.01, (err, d)
    # the module drop-in carries the option into its plan
    mk = lambda: modules.MambaBlocksSequential(1, bidirectional=True, d_model=64, fused_add_norm=False, rms_norm=True)
    net = modules.Dual_Path_Model_Skip(64, 64, mk(), mk(), num_layers=3, K=10, skip_around_intra=True, skip_n_block=1,
                                       linear_layer_after_inter_intra=False)
    net.load_state_dict(sds["masknet"], strict=True)
    mask = net.to(DEV)(g["mix_w"].to(DEV))
    assert rel_max(mask.cpu(), g["est_mask"]) <= 1e-3
    # and without the blend the result differs (the fixture really exercises it)
    est0 = DPSeparatorEngine(replace(hp, skip_n_block=0), sds, device=DEV, use_graph=False)(g["mix"].to(DEV)).cpu()
    assert rel_max(est0, g["est"]) > 1e-2


def test_dual_path_model_standalone_forward(golden_dir):
    path = os.path.join(golden_dir, "forward_dp_tiny_skip.npz")
    sds, g, _ = load_golden_forward(path)
    hp = DP_CONFIGS["tiny"]
    sep = modules.DPMambaSeparator.from_hparams(hp).load_reference_state_dicts(sds, strict=True).to(DEV)
    mask = sep.masknet(g["mix_w"].to(DEV))
    assert mask.shape == g["est_mask"].shape
    assert rel_max(mask.cpu(), g["est_mask"]) <= 1e-3


@pytest.mark.parametrize("name,B,T", [("XS", 2, 6000), ("S", 1, 8000), ("L", 1, 4000)])
def test_dpmamba_shipped_sizes_vs_oracle(name, B, T):
    """Shipped DPMamba hparams with K = 250 (intra sequences of 250 frames, inter sequences of S chunks) vs the oracle."""
    hp = replace(DP_CONFIGS[name], n_dp=2)          # 2 of the 8 / 16 identical-shape blocks keep the CPU oracle in seconds
    sds = init_dp_state_dicts(hp, 1234)
    mix, src = synth_mixture(B, T, seed=3)
    with torch.no_grad():
        ref = restate.separate_dp(mix, sds, hp, scan_impl="c")
    est = DPSeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False)(mix.to(DEV)).cpu()
    err, d = _gate(est, ref, src)
    print(f"dp_{name}: max-abs/rms {err:.3e} dSI-SNR {d:.2e}")
    assert err <= 1e-3 and d <= 0.01, (err, d)


def test_dpmamba_three_speakers_vs_oracle():
    hp = replace(DP_CONFIGS["XS"], n_dp=1, n_spk=3)
    sds = init_dp_state_dicts(hp, 5)
    mix, _ = synth_mixture(1, 4000, seed=2)
    with torch.no_grad():
        ref = restate.separate_dp(mix, sds, hp, scan_impl="c")
    est = DPSeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False)(mix.to(DEV)).cpu()
    assert est.shape == (1, 4000, 3)
    assert rel_max(est, ref) <= 1e-3


def test_dpmamba_bf16_mode_stated_tolerance():
    """bf16 mode (bf16 GEMM operands and xz / u / y storage; fp32 scan state, residual stream and GroupNorm statistics).
    Stated tolerance against the fp32 oracle, as for Mamba-TasNet: max-abs <= 0.15 * rms, SI-SNR(est, ref) >= 25 dB."""
    from avse_challenge_b200 import si_snr
    hp = replace(DP_CONFIGS["S"], n_dp=2)
    sds = init_dp_state_dicts(hp, 1234)
    mix, src = synth_mixture(1, 8000, seed=3)
    with torch.no_grad():
        ref = restate.separate_dp(mix, sds, hp, scan_impl="c")
    est = DPSeparatorEngine(hp, sds, device=DEV, mode="bf16", use_graph=False)(mix.to(DEV)).cpu()
    err, fid = rel_max(est, ref), si_snr(est, ref).min().item()
    print(f"dp_S bf16: max-abs/rms {err:.3e} SI-SNR(est,ref) {fid:.1f} dB")
    assert err <= 0.15 and fid >= 25.0, (err, fid)
    # and against the oracle that rounds where this mode rounds (restate.set_precision("product_bf16")).  Gate 2.5e-2 (six
    # times tighter than against fp32; Mamba-TasNet gets 1.5e-2): every dual block ends in a whole-utterance GroupNorm, so
    # a bf16 rounding flipped by an fp32-level difference (the 1-MUFU SiLU, accumulation order) reaches every sample
    restate.set_precision("product_bf16")
    try:
        with torch.no_grad():
            ref_m = restate.separate_dp(mix, sds, hp, scan_impl="c")
    finally:
        restate.set_precision("fp32")
    err_m, fid_m = rel_max(est, ref_m), si_snr(est, ref_m).min().item()
    print(f"dp_S bf16 vs matched-rounding oracle: max-abs/rms {err_m:.3e} SI-SNR {fid_m:.1f} dB")
    assert err_m <= 2.5e-2 and fid_m >= 40.0, (err_m, fid_m)


def test_dpmamba_unsupported_options_raise():
    mk = lambda **kw: modules.MambaBlocksSequential(1, bidirectional=True, d_model=64, fused_add_norm=False, rms_norm=True, **kw)
    with pytest.raises(NotImplementedError):
        modules.Dual_Path_Model(64, 64, mk(), mk(), linear_layer_after_inter_intra=True)
    with pytest.raises(NotImplementedError):
        modules.Dual_Path_Model(64, 64, mk(), mk(), linear_layer_after_inter_intra=False, use_global_pos_enc=True)
