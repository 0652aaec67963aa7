"""GPU (-m gpu): the one-launch streaming push (``mtn_stream_push_fwd``, one thread-block cluster per stream) against the
one-shot causal forward, the chunked batch plan and the CPU oracle.  It replaces, for chunks of <= 32 frames, the
reference's ``Mamba.step`` / ``inference_params`` streaming (``modules/mamba/bimamba.py:320-372``).
Gates: streaming == one-shot within 5e-5 x rms (two fp32-class summation orders of the same arithmetic); against the oracle
the north-star gate 1e-3 x rms, |dSI-SNR| <= 0.01 dB.
"""
import os

import pytest
import torch

from avse_challenge_b200 import CONFIGS, init_state_dicts, synth_mixture, pit_si_snr
from avse_challenge_b200 import _lib
from avse_challenge_b200.engine import SeparatorEngine
from avse_challenge_b200.streaming import StreamingSeparator
from oracle import restate
from tests.helpers import load_golden_forward, rel_max, hp_from_sds

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _stream(eng, mix, cuts, fused):
    st = StreamingSeparator(eng, mix.shape[0], use_graph=False, fused=fused)
    outs, pos = [], 0
    for c in cuts:
        outs.append(st.push(mix[:, pos:pos + c].contiguous()))
        pos += c
    assert pos == mix.shape[1]
    outs.append(st.flush())
    return torch.cat(outs, dim=1)


@pytest.mark.parametrize("chunk", [16, 8 * 3, 8 * 20, 8 * 32])
def test_fused_push_equals_one_shot_tiny(chunk):
    hp = CONFIGS["tiny"].causal()          # cluster of 2 CTAs, dt_rank 4
    sds = init_state_dicts(hp, 11)
    B, T = 2, 8 * 640
    mix, src = synth_mixture(B, T, seed=5)
    eng = SeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False)
    one = eng(mix.to(DEV)).cpu()
    first = chunk if chunk >= 16 else 16
    cuts = [first] + [chunk] * ((T - first) // chunk)
    cuts = cuts + ([T - sum(cuts)] if sum(cuts) < T else [])
    got = _stream(eng, mix.to(DEV), cuts, fused=True).cpu()
    assert got.shape == one.shape
    err = rel_max(got, one)
    print(f"fused push, chunk {chunk}: max-abs/rms vs one-shot {err:.3e}")
    assert err <= 5e-5, err
    with torch.no_grad():
        ref = restate.separate(mix, sds, hp.n_mamba, scan_impl="c")
    e2 = rel_max(got, ref)
    d2 = (pit_si_snr(got, src) - pit_si_snr(ref, src)).abs().max().item()
    assert e2 <= 1e-3 and d2 <= 0.01, (e2, d2)


def test_fused_and_batch_plan_share_one_stream_state():
    """Short pushes (cluster kernel) and long ones (batch plan) interleave on the same carried state."""
    hp = CONFIGS["tiny"].causal()
    sds = init_state_dicts(hp, 12)
    B = 3
    cuts = [160, 1000, 16, 8, 256, 2664, 64, 8 * 31]
    T = sum(cuts)
    mix, _ = synth_mixture(B, T, seed=6)
    eng = SeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False)
    one = eng(mix.to(DEV)).cpu()
    mixed = _stream(eng, mix.to(DEV), cuts, fused=None).cpu()
    plain = _stream(eng, mix.to(DEV), cuts, fused=False).cpu()
    e1, e2 = rel_max(mixed, one), rel_max(plain, one)
    print(f"mixed fused/batch pushes vs one-shot {e1:.3e}; batch plan only {e2:.3e}")
    assert e1 <= 5e-5 and e2 <= 5e-5


@pytest.mark.parametrize("name,B,frames", [("XS", 2, 20), ("S", 3, 20), ("S", 1, 7), ("L", 1, 24)])
def test_fused_push_shipped_sizes_vs_oracle(name, B, frames):
    """Cluster sizes 4 / 8 / 16 (the non-portable size needs the opt-in attribute), dt_rank 8 / 16 / 32."""
    hp = CONFIGS[name].causal()
    sds = init_state_dicts(hp, 1234)
    T = 8 * frames * 6 + 8
    mix, src = synth_mixture(B, T, seed=7)
    eng = SeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False)
    one = eng(mix.to(DEV)).cpu()
    cuts = [8 * frames + 8] + [8 * frames] * 5
    got = _stream(eng, mix.to(DEV), cuts, fused=True).cpu()
    err = rel_max(got, one)
    with torch.no_grad():
        ref = restate.separate(mix, sds, hp.n_mamba, scan_impl="c")
    e2 = rel_max(got, ref)
    d2 = (pit_si_snr(got, src) - pit_si_snr(ref, src)).abs().max().item()
    print(f"{hp.name} B={B} F={frames}: vs one-shot {err:.3e}, vs oracle {e2:.3e}, dSI-SNR {d2:.2e}")
    assert err <= 5e-5 and e2 <= 1e-3 and d2 <= 0.01, (err, e2, d2)


def test_fused_push_matches_reference_streaming_golden(golden_dir):
    """The reference's own inference_params run (prefill + step per token) of the tiny causal model."""
    sds, g, _ = load_golden_forward(os.path.join(golden_dir, "forward_tiny_causal.npz"))
    hp = hp_from_sds(sds).causal()
    eng = SeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False)
    mix = g["mix"]
    T = mix.shape[1] // 8 * 8
    cuts = [64] + [40] * ((T - 64) // 40)
    cuts += [T - sum(cuts)] if sum(cuts) < T else []
    got = _stream(eng, mix[:, :T].contiguous().to(DEV), cuts, fused=True).cpu()
    ref = g["est"][:, :got.shape[1]]
    err = rel_max(got[:, :ref.shape[1]], ref)
    print(f"fused push vs reference causal golden: {err:.3e}")
    assert err <= 1e-3, err


def test_fused_push_rejects_what_it_does_not_implement():
    hp = CONFIGS["tiny"].causal()
    sds = init_state_dicts(hp, 3)
    eng = SeparatorEngine(hp, sds, device=DEV, mode="bf16", use_graph=False)
    with pytest.raises(_lib.MtnError):
        StreamingSeparator(eng, 1, fused=True)
    assert StreamingSeparator(eng, 1)._fused is None      # auto: falls back to the batch plan's chunk kernels


@pytest.mark.parametrize("name,B", [("tiny", 3), ("S", 2)])
def test_stack_decode_calls_take_the_one_launch_kernel_and_share_the_reference_caches(name, B):
    """`MambaBlocksSequential.forward(x, inference_params)` (modules/mamba_blocks.py:186-197 over bimamba.py:320-372): calls of
    <= 32 tokens run as one cluster-kernel launch on the reference's own cache tensors (kv[i] = (conv_state [B, di, 4],
    ssm_state [B, di, 16]), here views of two stacked buffers); longer calls take the chunk kernels on the same caches.  Any
    split of the sequence reproduces the one-call forward."""
    import types
    from avse_challenge_b200 import modules
    hp = CONFIGS[name].causal()
    m = init_state_dicts(hp, 5)["masknet"]
    net = modules.MambaBlocksSequential(hp.n_mamba, bidirectional=False, d_model=hp.d_model, fused_add_norm=False, rms_norm=True)
    net.load_state_dict({k[len("mamba_net."):]: v for k, v in m.items() if k.startswith("mamba_net.")}, strict=True)
    net.to(DEV)
    g = torch.Generator().manual_seed(9)
    cuts = [40, 1, 1, 7, 32, 50, 1, 20]          # 40 and 50: chunk kernels; the rest: the one-launch kernel
    x = torch.randn(B, sum(cuts), hp.d_model, generator=g).to(DEV)
    full = net(x)
    ip = types.SimpleNamespace(seqlen_offset=0, key_value_memory_dict={})
    outs, pos = [], 0
    for c in cuts:
        outs.append(net(x[:, pos:pos + c], inference_params=ip))
        pos += c
        ip.seqlen_offset = pos
    got = torch.cat(outs, dim=1)
    err = rel_max(got.cpu(), full.cpu())
    print(f"{hp.name}: split decode vs one call {err:.3e}")
    assert err <= 5e-5, err
    kv = ip.key_value_memory_dict
    cs, ss = kv[hp.n_mamba - 1]
    assert tuple(cs.shape) == (B, hp.d_inner, 4) and tuple(ss.shape) == (B, hp.d_inner, 16)
    assert cs.data_ptr() == kv["_mtn_b200"]["conv"][hp.n_mamba - 1].data_ptr()      # views of the stacked buffers
    # caches a caller allocated itself (Mamba.allocate_inference_cache) are adopted, not ignored
    ip2 = types.SimpleNamespace(seqlen_offset=pos - 20, key_value_memory_dict={})
    ipa = types.SimpleNamespace(seqlen_offset=0, key_value_memory_dict={})
    net(x[:, :pos - 20], inference_params=ipa)                                      # long call: chunk kernels, plain tensors
    for i in range(hp.n_mamba):
        ip2.key_value_memory_dict[i] = tuple(t.clone() for t in ipa.key_value_memory_dict[i])
    last = net(x[:, pos - 20:], inference_params=ip2)
    assert rel_max(last.cpu(), full[:, pos - 20:].cpu()) <= 5e-5


@pytest.mark.parametrize("name,B,frames,dsl", [("tiny", 2, 20, 32), ("tiny", 2, 9, 64), ("XS", 2, 20, 32), ("XS", 1, 32, 128),
                                               ("S", 2, 20, 32), ("S", 3, 20, 128), ("S", 1, 2, 32), ("L", 1, 12, 128)])
def test_fused_push_channels_per_cta_variants(name, B, frames, dsl):
    """The same kernel cut 32 / 64 / 128 d_inner channels per CTA (cluster size 2 * d_model / dsl: 16-CTA clusters for the
    lowest latency, 4-CTA clusters to keep many streams resident): each equals the one-shot forward and the oracle."""
    hp = CONFIGS[name].causal()
    sds = init_state_dicts(hp, 1234)
    T = 8 * frames * 5 + 8
    mix, src = synth_mixture(B, T, seed=7)
    eng = SeparatorEngine(hp, sds, device=DEV, mode="fp32", use_graph=False)
    one = eng(mix.to(DEV)).cpu()
    st = StreamingSeparator(eng, B, use_graph=False, fused=True, channels_per_cta=dsl)
    outs, pos = [], 0
    for c in [8 * frames + 8] + [8 * frames] * 4:
        outs.append(st.push(mix[:, pos:pos + c].contiguous().to(DEV)))
        pos += c
    outs.append(st.flush())
    got = torch.cat(outs, dim=1).cpu()
    err = rel_max(got, one)
    with torch.no_grad():
        ref = restate.separate(mix, sds, hp.n_mamba, scan_impl="c")
    e2 = rel_max(got, ref)
    print(f"{hp.name} dsl={dsl} B={B} F={frames}: vs one-shot {err:.3e}, vs oracle {e2:.3e}")
    assert err <= 5e-5 and e2 <= 1e-3, (err, e2)


def test_channels_per_cta_policy():
    from avse_challenge_b200.stream_fused import channels_per_cta
    assert channels_per_cta(1, 256) == 32 and channels_per_cta(4, 256) == 32       # 16-CTA clusters while all of them are resident
    assert channels_per_cta(8, 256) == 64 and channels_per_cta(12, 256) == 64
    assert channels_per_cta(16, 256) == 128 and channels_per_cta(32, 256) == 128 and channels_per_cta(64, 256) == 128
    assert channels_per_cta(1, 512) == 64                                          # d_model 512: 32 would need 32 CTAs
    assert channels_per_cta(32, 512, 32) == 64 and channels_per_cta(32, 512, 8) == 128   # shared memory decides at 32 frames
    assert channels_per_cta(1, 64) == 32 and channels_per_cta(64, 64) == 64        # d_model 64: 128 would span both speakers
