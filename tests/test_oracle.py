"""CPU: the oracle restatement is pinned against outputs of the REAL reference (tests/golden/*.npz, minted by
oracle/make_golden.py from /root/reference under leaf shims -- the reference ships no tests of its own)."""
import os

import numpy as np
import pytest
import torch

from oracle import restate
from tests.helpers import load_golden_forward, rel_max, hp_from_sds


def test_scan_restatement_matches_reference_selective_scan_ref(golden_dir):
    z = np.load(os.path.join(golden_dir, "scan_ref.npz"))
    t = lambda k: torch.from_numpy(z[k])
    # reference layout [B, D, L] / [B, N, L]  ->  channel-last
    u, delta, zz = (t(k).transpose(1, 2).contiguous() for k in ("u", "delta", "z"))
    Bm, Cm = (t(k).transpose(1, 2).contiguous() for k in ("B", "C"))
    for impl in ("torch", "c"):
        out, last = restate.selective_scan(u, delta, t("A"), Bm, Cm, t("D"), zz, t("bias"), impl=impl)
        assert rel_max(out.transpose(1, 2), t("out")) < 2e-6, impl
        assert rel_max(last, t("last_state")) < 2e-6, impl
        out_ng, _ = restate.selective_scan(u, delta, t("A"), Bm, Cm, t("D"), None, t("bias"), impl=impl, gate=False)
        assert rel_max(out_ng.transpose(1, 2), t("out_nogate")) < 2e-6, impl
        out_r, _ = restate.selective_scan(u, delta, t("A"), Bm, Cm, t("D"), zz, t("bias"), reverse=True, impl=impl)
        assert rel_max(out_r.transpose(1, 2), t("out_reverse")) < 2e-6, impl


def test_scan_chunk_carry_equals_one_shot(golden_dir):
    z = np.load(os.path.join(golden_dir, "scan_ref.npz"))
    t = lambda k: torch.from_numpy(z[k])
    u, delta, zz = (t(k).transpose(1, 2).contiguous() for k in ("u", "delta", "z"))
    Bm, Cm = (t(k).transpose(1, 2).contiguous() for k in ("B", "C"))
    full, last = restate.selective_scan(u, delta, t("A"), Bm, Cm, t("D"), zz, t("bias"), impl="c")
    cut = 61
    a, ha = restate.selective_scan(u[:, :cut], delta[:, :cut], t("A"), Bm[:, :cut], Cm[:, :cut], t("D"), zz[:, :cut],
                                   t("bias"), impl="c")
    b, hb = restate.selective_scan(u[:, cut:], delta[:, cut:], t("A"), Bm[:, cut:], Cm[:, cut:], t("D"), zz[:, cut:],
                                   t("bias"), impl="c", h_in=ha)
    assert torch.equal(torch.cat([a, b], 1), full)
    assert torch.equal(hb, last)


@pytest.mark.parametrize("tag", ["tiny_refinit", "tiny_trained"])
def test_forward_restatement_matches_reference(golden_dir, tag):
    sds, g, taps = load_golden_forward(os.path.join(golden_dir, f"forward_{tag}.npz"))
    hp = hp_from_sds(sds)
    my_taps = []
    with torch.no_grad():
        est = restate.separate(g["mix"], sds, hp.n_mamba, scan_impl="c", taps=my_taps)
    assert est.shape == g["est"].shape
    assert rel_max(est, g["est"]) < 1e-5
    mask, mix_w = my_taps[-1], my_taps[-2]
    assert rel_max(mix_w.transpose(1, 2), g["mix_w"]) < 1e-5           # reference layout [B, N, L]
    assert rel_max(mask.permute(0, 1, 3, 2), g["est_mask"]) < 1e-5     # reference layout [spk, B, N, L]
    assert rel_max(my_taps[0], taps["mixer0_out"]) < 1e-5


def _match_stats(a, ref):
    d = (a.double() - ref.double()).abs()
    rms = ref.double().pow(2).mean().sqrt()
    return (d == 0).double().mean().item(), (d.mean() / rms).item(), (d.max() / rms).item()


def test_autocast_restatement_matches_reference_autocast_run(golden_dir):
    """``set_precision("autocast_ref")`` is pinned against the reference's own modules run under
    ``torch.autocast(bfloat16)`` (``forward_tiny_autocast.npz``, minted by oracle/make_golden.py; ``precision: bf16`` is
    the recipes' default, mambatasnet_S.yaml:38).  A rounding model that rounds at the same points reproduces bf16 results
    bit for bit except where fp32 summation order flips a rounding: nearly all elements must be identical."""
    sds, g, taps = load_golden_forward(os.path.join(golden_dir, "forward_tiny_autocast.npz"))
    hp = hp_from_sds(sds)
    restate.set_precision("autocast_ref")
    try:
        my_taps = []
        with torch.no_grad():
            est = restate.separate(g["mix"], sds, hp.n_mamba, scan_impl="c", taps=my_taps)
    finally:
        restate.set_precision("fp32")
    mask, mix_w = my_taps[-1], my_taps[-2]
    assert torch.equal(mix_w.transpose(1, 2), g["mix_w"])                      # encoder: no summation-order freedom
    for name, a, ref, min_exact in (("mixer0", my_taps[0], taps["mixer0_out"], 0.995),
                                    ("mask", mask.permute(0, 1, 3, 2), g["est_mask"], 0.995),
                                    ("est", est, g["est"], 0.98)):
        exact, mean_err, max_err = _match_stats(a, ref)
        assert exact >= min_exact and mean_err <= 2e-4 and max_err <= 3e-2, (name, exact, mean_err, max_err)
    # the fp32 restatement is far from this run: the rounding points, not noise, are what the model adds
    with torch.no_grad():
        est32 = restate.separate(g["mix"], sds, hp.n_mamba, scan_impl="c")
    assert _match_stats(est32, g["est"])[1] > 20 * _match_stats(est, g["est"])[1]


def test_precision_models_order_of_error():
    """On the same input the product's bf16 rounding model (fp32 residual stream, fp32 [dt|B|C] and delta) is closer to the
    fp32 reference than the reference's own autocast path (bf16 residual stream and delta) -- the numbers DESIGN.md 2
    quotes; also checks that set_precision restores cleanly."""
    from avse_challenge_b200 import CONFIGS, init_state_dicts, synth_mixture
    hp = CONFIGS["tiny"]
    sds = init_state_dicts(hp, 11)
    mix, _ = synth_mixture(2, 1603, seed=3)
    out = {}
    for mode in ("fp32", "product_bf16", "autocast_ref", "fp32"):
        restate.set_precision(mode)
        with torch.no_grad():
            cur = restate.separate(mix, sds, hp.n_mamba, scan_impl="c")
        if mode in out:
            assert torch.equal(cur, out[mode])                                  # the switch leaves no state behind
        out[mode] = cur
    e_prod = _match_stats(out["product_bf16"], out["fp32"])[1]
    e_ref = _match_stats(out["autocast_ref"], out["fp32"])[1]
    assert 0 < e_prod < e_ref < 0.05, (e_prod, e_ref)


def test_torch_loop_scan_equals_c_scan(golden_dir):
    sds, g, _ = load_golden_forward(os.path.join(golden_dir, "forward_tiny_trained.npz"))
    hp = hp_from_sds(sds)
    with torch.no_grad():
        a = restate.separate(g["mix"][:1, :400], sds, hp.n_mamba, scan_impl="torch")
        b = restate.separate(g["mix"][:1, :400], sds, hp.n_mamba, scan_impl="c")
    assert rel_max(a, b) < 1e-5


def test_gemm_rounding_models_are_ordered(golden_dir):
    """bf16x3 (what the CUDA fp32 mode computes) must sit far inside the 1e-3 gate; plain tf32/bf16 do not."""
    sds, g, _ = load_golden_forward(os.path.join(golden_dir, "forward_tiny_trained.npz"))
    hp = hp_from_sds(sds)
    errs = {}
    try:
        with torch.no_grad():
            ref = restate.separate(g["mix"], sds, hp.n_mamba)
            for mode in ("bf16x3", "tf32", "bf16"):
                restate.set_gemm_mode(mode)
                errs[mode] = rel_max(restate.separate(g["mix"], sds, hp.n_mamba), ref)
    finally:
        restate.set_gemm_mode("fp32")
    assert errs["bf16x3"] < 1e-4 < errs["tf32"] < errs["bf16"]


def test_si_snr_restatement_matches_reference_cal_si_snr(golden_dir):
    """oracle.restate.cal_si_snr against outputs of the reference's own cal_si_snr (baseline/avse2/utils/dnn.py:15-57),
    minted by oracle/make_golden.py::golden_si_snr."""
    z = np.load(os.path.join(golden_dir, "si_snr_ref.npz"))
    src, est, ref = torch.from_numpy(z["src"]), torch.from_numpy(z["est"]), torch.from_numpy(z["si_snr"]).double()
    mine = restate.cal_si_snr(src, est)
    assert (mine - ref).abs().max() < 1e-3          # the reference computes in fp32; dB
    best, imp, perm, pairs = restate.pit_si_snr_improvement(est, src, src.sum(-1))
    assert perm.tolist() == [1, 0, 0]               # utterance 0 was minted with its speakers swapped
    assert torch.allclose(pairs[1:, [0, 1], [0, 1]], mine[1:], atol=1e-9)
    assert (best > 4).all() and (imp > 4).all() and best[1] > 40


# ------------------------------------------------------------------------------------------------ causal / streaming
def test_causal_restatement_matches_reference_golden(golden_dir):
    """bidirectional=False: the reference's own run (vendored unidirectional branch, oracle/ref_shims.UniMamba)."""
    sds, out, taps = load_golden_forward(os.path.join(golden_dir, "forward_tiny_causal.npz"))
    assert not any(k.endswith("A_b_log") for k in sds["masknet"])
    n = 1 + max(int(k.split(".")[2]) for k in sds["masknet"] if k.startswith("mamba_net.layers."))
    for impl in ("torch", "c"):
        est = restate.separate(out["mix"], sds, n, scan_impl=impl)
        assert (est - out["est"]).abs().max().item() <= 2e-6, impl


def test_streaming_restatement_matches_reference_inference_cache(golden_dir):
    """Prefill + Mamba.step through the reference's inference_params caches (golden) vs the oracle's chunk-carry
    restatement, for two different chunkings, including the literal per-token `mixer_step`."""
    sds, _, _ = load_golden_forward(os.path.join(golden_dir, "forward_tiny_causal.npz"))
    z = np.load(os.path.join(golden_dir, "stream_tiny_causal.npz"))
    h = torch.from_numpy(z["h"])
    m = sds["masknet"]
    n = 1 + max(int(k.split(".")[2]) for k in m if k.startswith("mamba_net.layers."))
    L0 = int(z["L0"])
    for cuts in ([L0] + [1] * (h.shape[1] - L0), [5, 2, 1, 13, 20]):
        st = restate.new_stream_state(m, n, h.shape[0])
        outs, pos = [], 0
        for c in cuts:
            outs.append(restate.mamba_stack_fwd(h[:, pos:pos + c], m, n, states=st))
            pos += c
        assert pos == h.shape[1]
        got = torch.cat(outs, dim=1)
        assert (got - torch.from_numpy(z["streamed"])).abs().max().item() <= 2e-6
        assert (got - torch.from_numpy(z["full"])).abs().max().item() <= 2e-6
        for i in range(n):
            assert (st[i]["conv"] - torch.from_numpy(z[f"conv_state/{i}"])).abs().max().item() <= 1e-6
            assert (st[i]["ssm"] - torch.from_numpy(z[f"ssm_state/{i}"])).abs().max().item() <= 1e-6
    # literal step: layer 0 mixer, one token from zero caches == first output row of the chunk form
    p = "mamba_net.layers.0.mixer."
    cs, ss = torch.zeros(h.shape[0], m[p + "A_log"].shape[0], 4), torch.zeros(h.shape[0], m[p + "A_log"].shape[0], 16)
    st0 = {"conv": cs.clone(), "ssm": ss.clone()}
    x = h[:, :3]
    chunk = restate.mixer_fwd(x, m, p, state=st0)
    rows = [restate.mixer_step(x[:, t], m, p, cs, ss) for t in range(3)]
    assert (torch.stack(rows, dim=1) - chunk).abs().max().item() <= 2e-6
    assert (cs - st0["conv"]).abs().max().item() <= 1e-6 and (ss - st0["ssm"]).abs().max().item() <= 1e-6


# ------------------------------------------------------------------------------------------------ DPMamba
@pytest.mark.parametrize("tag", ["dp_tiny_skip", "dp_tiny_noskip", "dp_tiny_blockskip"])
def test_dpmamba_restatement_matches_reference_golden(golden_dir, tag):
    """Golden = the reference's vendored Dual_Path_Model_Skip.forward over the real MambaBlocksSequential stacks."""
    from dataclasses import replace
    from avse_challenge_b200.hparams import DP_CONFIGS
    path = os.path.join(golden_dir, f"forward_{tag}.npz")
    sds, out, _ = load_golden_forward(path)
    z = np.load(path)
    hp = replace(DP_CONFIGS["tiny"], skip_around_intra=bool(z["skip_around_intra"]), chunk_size=int(z["chunk_size"]),
                 n_dp=int(z["n_dp"]), skip_n_block=int(z["skip_n_block"]) if "skip_n_block" in z.files else 0)
    with torch.no_grad():
        mix_w = restate.encoder_fwd(out["mix"], sds["encoder"]["conv1d.weight"])
        mask = restate.dp_masknet_fwd(mix_w, sds["masknet"], hp.n_dp, hp.chunk_size, hp.skip_around_intra, scan_impl="c",
                                      skip_n_block=hp.skip_n_block)
        est = restate.separate_dp(out["mix"], sds, hp, scan_impl="c")
    assert (mask.permute(0, 1, 3, 2) - out["est_mask"]).abs().max().item() <= 2e-6
    assert (est - out["est"]).abs().max().item() <= 2e-6


@pytest.mark.parametrize("L,K", [(3999, 250), (7, 10), (10, 10), (249, 250), (250, 250), (375, 250), (11999, 250)])
def test_dp_segmentation_round_trip(L, K):
    """over_add(segment(x)) == 2 x (every frame is covered by exactly two chunks), chunk count as the reference's."""
    x = torch.randn(2, L, 3)
    seg, gap = restate.dp_segment(x, K)
    S, gap2 = restate.dp_num_chunks(L, K)
    assert seg.shape == (2, S, K, 3) and gap == gap2 and S % 2 == 0
    assert torch.allclose(restate.dp_over_add(seg, gap), 2 * x)


def test_softmax_mask_restatement_matches_reference_golden(golden_dir):
    """mask_nonlinear="softmax" (mamba_masknet.py:133-134: softmax over dim 2 of [spk, B, N, L] = the N channels)."""
    sds, out, _ = load_golden_forward(os.path.join(golden_dir, "forward_tiny_softmax.npz"))
    with torch.no_grad():
        mix_w = restate.encoder_fwd(out["mix"], sds["encoder"]["conv1d.weight"])
        mask = restate.masknet_fwd(mix_w, sds["masknet"], 2, scan_impl="c", mask_nonlinear="softmax")
        est = restate.separate(out["mix"], sds, 2, scan_impl="c", mask_nonlinear="softmax")
    assert (mask.permute(0, 1, 3, 2) - out["est_mask"]).abs().max().item() <= 2e-6
    assert torch.allclose(out["est_mask"].sum(dim=2), torch.ones(2, 1, out["est_mask"].shape[3]), atol=1e-5)
    assert (est - out["est"]).abs().max().item() <= 2e-6


def test_layernorm_blocks_restatement_matches_reference_golden(golden_dir):
    """rms_norm=False: nn.LayerNorm blocks and norm_f (modules/mamba_blocks.py:36-41,167-169), reference's own run."""
    sds, out, _ = load_golden_forward(os.path.join(golden_dir, "forward_tiny_layernorm.npz"))
    assert "mamba_net.norm_f.bias" in sds["masknet"]
    with torch.no_grad():
        est = restate.separate(out["mix"], sds, 2, scan_impl="c")
    assert (est - out["est"]).abs().max().item() <= 2e-6
