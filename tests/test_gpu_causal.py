"""GPU (-m gpu): the unidirectional / causal separator and its streaming forms against the golden
vectors minted from the reference's own unidirectional branch + inference caches (oracle/make_golden.py) and against
the CPU oracle.  Gates as in test_gpu_parity.py: max|est - ref| <= 1e-3 * rms(ref), |dSI-SNR| <= 0.01 dB.
"""
import os
import types

import numpy as np
import pytest
import torch

from avse_challenge_b200 import CONFIGS, init_state_dicts, synth_mixture, pit_si_snr
from avse_challenge_b200 import _lib, ops, modules
from avse_challenge_b200.engine import SeparatorEngine, MambaStack
from avse_challenge_b200.streaming import StreamingSeparator
from oracle import restate
from tests.helpers import load_golden_forward, rel_max, hp_from_sds

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _gate(est, ref, src):
    err = rel_max(est, ref)
    d = (pit_si_snr(est, src) - pit_si_snr(ref, src)).abs().max().item()
    return err, d


def _causal_hp(sds):
    hp = hp_from_sds(sds)
    return hp.causal()


def test_conv_dir_mask_forward_only():
    g = torch.Generator().manual_seed(3)
    B, L, di = 2, 77, 128
    xz = torch.randn(B * L, 2 * di, generator=g).to(DEV)
    w = torch.randn(2, di, 4, generator=g).to(DEV)
    b = torch.randn(2, di, generator=g).to(DEV)
    both = ops.conv_silu(xz, w, b, B, L, di, 2)
    u = torch.full_like(both, 7.0)
    ops.conv_silu(xz, w[:1].contiguous(), b[:1].contiguous(), B, L, di, 2, u=u, dir_mask=1)
    assert torch.equal(u[:, :, :di], both[:, :, :di])
    assert (u[:, :, di:] == 7.0).all()          # backward half untouched
    with pytest.raises(_lib.MtnError):
        ops.conv_silu(xz, w, b, B, L, di, 2, dir_mask=2)


@pytest.mark.parametrize("cuts", [[40], [1, 1, 38], [7, 20, 13]])
def test_decoder_stream_tail_equals_one_shot(cuts):
    g = torch.Generator().manual_seed(4)
    B, L, N = 3, 40, 64
    sep = torch.randn(B, L, 2 * N, generator=g).to(DEV)
    w = torch.randn(N, 16, generator=g).to(DEV)
    T = (L - 1) * 8 + 16
    one = ops.decoder(sep.view(B * L, 2 * N), w, B, T, L, N)
    tail = torch.zeros(B, 2, 8, device=DEV)
    outs, pos = [], 0
    for c in cuts:
        chunk = sep[:, pos:pos + c].contiguous().view(B * c, 2 * N)
        outs.append(ops.decoder(chunk, w, B, 8 * c, c, N, tail=tail).clone())
        pos += c
    outs.append(tail.transpose(1, 2))
    assert torch.equal(torch.cat(outs, dim=1), one)     # same additions in the same order: bit-exact
    with pytest.raises(_lib.MtnError):
        ops.decoder(sep[:, :4].contiguous().view(B * 4, 2 * N), w, B, 40, 4, N, tail=tail)


@pytest.mark.parametrize("use_graph", [False, True])
def test_causal_end_to_end_matches_reference_golden(golden_dir, use_graph):
    sds, g, _ = load_golden_forward(os.path.join(golden_dir, "forward_tiny_causal.npz"))
    hp = _causal_hp(sds)
    sep = modules.MambaTasNetSeparator.from_hparams(hp, mode="fp32", use_graph=use_graph)
    sep.load_reference_state_dicts(sds, strict=True).to(DEV)
    est = sep(g["mix"].to(DEV)).cpu()
    err, d = _gate(est, g["est"], g["src"])
    assert err <= 1e-3 and d <= 0.01, (err, d)


@pytest.mark.parametrize("name,B,T", [("S", 2, 16000), ("L", 1, 4000)])
def test_causal_shipped_sizes_vs_oracle(name, B, T):
    hp = CONFIGS[name].causal()
    sds = init_state_dicts(hp, 1234)
    mix, src = synth_mixture(B, T, seed=7)
    with torch.no_grad():
        ref = restate.separate(mix, sds, hp.n_mamba, scan_impl="c")
    est = SeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False)(mix.to(DEV)).cpu()
    err, d = _gate(est, ref, src)
    print(f"{hp.name}: max-abs/rms {err:.3e} dSI-SNR {d:.2e}")
    assert err <= 1e-3 and d <= 0.01, (err, d)


def test_causal_output_does_not_depend_on_the_future():
    hp = CONFIGS["tiny"].causal()
    sds = init_state_dicts(hp, 9)
    mix, _ = synth_mixture(1, 4000, seed=3)
    eng = SeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False)
    a = eng(mix.to(DEV))
    mix2 = mix.clone()
    mix2[:, 2400:] = torch.randn(1, 1600) * 0.1
    b = eng(mix2.to(DEV))
    # frame l covers input samples [8l, 8l+16) and writes output samples [8l, 8l+16): frame 299 is the first to see a
    # changed sample (2400..2407), so outputs before 8*299 = 2392 must be bit-identical
    assert torch.equal(a[:, :2392], b[:, :2392])
    assert not torch.equal(a[:, 2392:2400], b[:, 2392:2400])


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("chunk", [16, 8 * 20, 8 * 125, 8 * 333])
def test_streaming_equals_one_shot(chunk, use_graph):
    """Chunked streaming (carried encoder overlap, conv history, SSM state, overlap-add tail) reproduces the one-shot
    causal forward, down to chunks of two frames; the first push must hold one full encoder window."""
    hp = CONFIGS["tiny"].causal()
    sds = init_state_dicts(hp, 11)
    B, T = 2, 8 * 1000 + 5
    mix, src = synth_mixture(B, T, seed=5)
    eng = SeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False)
    one = eng(mix.to(DEV)).cpu()
    st = StreamingSeparator(eng, B, use_graph=use_graph)
    got = st.separate(mix.to(DEV), chunk).cpu()
    assert got.shape == one.shape
    err = rel_max(got, one)
    print(f"chunk {chunk}: max-abs/rms vs one-shot {err:.3e}")
    assert err <= 2e-5, err
    # and the one-shot itself is within the gate of the oracle, so streaming is too
    with torch.no_grad():
        ref = restate.separate(mix, sds, hp.n_mamba, scan_impl="c")
    e2, d2 = _gate(got, ref, src)
    assert e2 <= 1e-3 and d2 <= 0.01


@pytest.mark.parametrize("chunk", [16, 8 * 20, 8 * 333])
def test_streaming_fused_norm_plan_equals_one_shot(chunk):
    """The plan with Add -> RMSNorm folded into the GEMM epilogues (17 fewer launches per push) streams to the same
    samples as the one-shot forward of the default plan."""
    hp = CONFIGS["tiny"].causal()
    sds = init_state_dicts(hp, 11)
    B, T = 2, 8 * 1000 + 5
    mix, _ = synth_mixture(B, T, seed=5)
    one = SeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False)(mix.to(DEV)).cpu()
    eng = SeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False, fuse_norm=True)
    got = StreamingSeparator(eng, B, use_graph=True).separate(mix.to(DEV), chunk).cpu()
    err = rel_max(got, one)
    print(f"fused-norm streaming, chunk {chunk}: max-abs/rms vs one-shot {err:.3e}")
    assert err <= 2e-4, err


def test_streaming_single_frame_chunks_and_rejects():
    hp = CONFIGS["tiny"].causal()
    sds = init_state_dicts(hp, 12)
    mix, _ = synth_mixture(1, 8 * 60, seed=6)
    eng = SeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False)
    one = eng(mix.to(DEV)).cpu()
    st = StreamingSeparator(eng, 1, use_graph=False)
    outs = [st.push(mix[:, :16].to(DEV))]
    for k in range(2, 60):
        outs.append(st.push(mix[:, 8 * k: 8 * k + 8].to(DEV)))     # one hop = one new frame per call
    outs.append(st.flush())
    got = torch.cat(outs, dim=1).cpu()
    assert rel_max(got, one) <= 2e-5
    with pytest.raises(_lib.MtnError):
        st.push(mix[:, :12].to(DEV))
    with pytest.raises(NotImplementedError):
        StreamingSeparator(SeparatorEngine(CONFIGS["tiny"], init_state_dicts(CONFIGS["tiny"], 1), device=DEV), 1)


def test_causal_bf16_mode_one_shot_and_streaming():
    """bf16 mode of the causal stack: stated tolerance against the fp32 oracle (as test_bf16_mode_stated_tolerance), and
    chunked streaming against the one-shot bf16 forward (same kernels, same rounding points -> fp32-class agreement)."""
    from avse_challenge_b200 import si_snr
    hp = CONFIGS["S"].causal()
    sds = init_state_dicts(hp, 1234)
    mix, src = synth_mixture(2, 8000, seed=7)
    with torch.no_grad():
        ref = restate.separate(mix, sds, hp.n_mamba, scan_impl="c")
    eng = SeparatorEngine(hp, sds, device=DEV, mode="bf16", use_graph=False)
    one = eng(mix.to(DEV)).cpu()
    err, fid = rel_max(one, ref), si_snr(one, ref).min().item()
    print(f"causal S bf16: max-abs/rms {err:.3e}  SI-SNR(est,ref) {fid:.1f} dB")
    assert err <= 0.15 and fid >= 25.0, (err, fid)
    got = StreamingSeparator(eng, 2, use_graph=True).separate(mix.to(DEV), 8 * 40).cpu()
    assert rel_max(got, one) <= 1e-4


def test_stack_forward_with_inference_params_matches_reference_cache_golden(golden_dir):
    """`MambaBlocksSequential.forward(x, inference_params)` drop-in: prefill L0 tokens then one token per call, against
    the reference's own streamed outputs and final conv / ssm caches."""
    sds, _, _ = load_golden_forward(os.path.join(golden_dir, "forward_tiny_causal.npz"))
    z = np.load(os.path.join(golden_dir, "stream_tiny_causal.npz"))
    hp = _causal_hp(sds)
    net = modules.MambaBlocksSequential(hp.n_mamba, bidirectional=False, d_model=hp.d_model, fused_add_norm=False,
                                        rms_norm=True)
    net.load_state_dict({k[len("mamba_net."):]: v for k, v in sds["masknet"].items() if k.startswith("mamba_net.")},
                        strict=True)
    net.to(DEV)
    h = torch.from_numpy(z["h"]).to(DEV)
    full = net(h).cpu()
    assert rel_max(full, torch.from_numpy(z["full"])) <= 1e-4
    ip = types.SimpleNamespace(seqlen_offset=0, key_value_memory_dict={})
    L0 = int(z["L0"])
    outs = [net(h[:, :L0], inference_params=ip)]
    ip.seqlen_offset = L0
    for t in range(L0, h.shape[1]):
        outs.append(net(h[:, t:t + 1], inference_params=ip))
        ip.seqlen_offset += 1
    got = torch.cat(outs, dim=1).cpu()
    assert rel_max(got, torch.from_numpy(z["streamed"])) <= 1e-4
    for i in range(hp.n_mamba):
        cs, ss = ip.key_value_memory_dict[i]
        assert rel_max(cs.cpu(), torch.from_numpy(z[f"conv_state/{i}"])) <= 1e-4
        assert rel_max(ss.cpu(), torch.from_numpy(z[f"ssm_state/{i}"])) <= 1e-4
    # the reference's reset idiom: the SAME cache object with seqlen_offset = 0 starts a new sequence from zero state
    # (prefill overwrites the caches, bimamba.py:271-304) -- nothing of the previous sequence may leak
    ip.seqlen_offset = 0
    again = net(h[:, :L0], inference_params=ip).cpu()
    assert torch.equal(again, outs[0].cpu())


def test_mixer_step_matches_oracle_step():
    """`Mamba.step` drop-in (bimamba.py:320-372) vs the literal restatement, several tokens, caches compared."""
    hp = CONFIGS["tiny"].causal()
    m = init_state_dicts(hp, 21)["masknet"]
    p = "mamba_net.layers.1.mixer."
    mixer = modules.Mamba(hp.d_model, bimamba_type="none")
    mixer.load_state_dict({k[len(p):]: v for k, v in m.items() if k.startswith(p)}, strict=True)
    mixer.to(DEV)
    B = 3
    g = torch.Generator().manual_seed(8)
    xs = torch.randn(B, 6, hp.d_model, generator=g)
    cs_ref, ss_ref = torch.zeros(B, hp.d_inner, 4), torch.zeros(B, hp.d_inner, 16)
    cs, ss = mixer.allocate_inference_cache(B)
    for t in range(6):
        ref = restate.mixer_step(xs[:, t], m, p, cs_ref, ss_ref)
        out, cs, ss = mixer.step(xs[:, t:t + 1].to(DEV), cs, ss)
        assert out.shape == (B, 1, hp.d_model)
        assert rel_max(out[:, 0].cpu(), ref) <= 1e-4, t
        assert rel_max(cs.cpu(), cs_ref) <= 1e-4 and rel_max(ss.cpu(), ss_ref) <= 1e-4, t


def test_stack_forward_bidirectional_matches_oracle():
    """The stand-alone stack on many short sequences (the DPMamba intra / inter shapes: L = 250 and L = 34)."""
    hp = CONFIGS["tiny"]
    m = init_state_dicts(hp, 31)["masknet"]
    sd = {k[len("mamba_net."):]: v for k, v in m.items() if k.startswith("mamba_net.")}
    stack = MambaStack(hp, sd, device=DEV, mode="fp32")
    g = torch.Generator().manual_seed(2)
    for Bs, L in ((9, 250), (40, 34), (5, 1)):
        x = torch.randn(Bs, L, hp.d_model, generator=g)
        with torch.no_grad():
            ref = restate.mamba_stack_fwd(x, m, hp.n_mamba, scan_impl="c")
        got = stack(x.to(DEV)).cpu()
        assert rel_max(got, ref) <= 1e-4, (Bs, L, rel_max(got, ref))
