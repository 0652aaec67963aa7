"""TEST INFRASTRUCTURE ONLY -- mint golden vectors from the REAL reference (build container only).

Runs the reference's own modules (imported from ``/root/reference`` under ``oracle/ref_shims.py``)
on seeded inputs and stores inputs, weights and outputs as small ``.npz`` fixtures under
``tests/golden/``.  The reference ships no tests or golden vectors for this path (SURVEY.md
section 4), so these files are what pins both the oracle restatement and the CUDA path.

    python -m oracle.make_golden
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from avse_challenge_b200.hparams import CONFIGS, DP_CONFIGS, init_state_dicts, init_dp_state_dicts  # noqa: E402
from avse_challenge_b200.synth import synth_mixture  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(sd, prefix):
    return {f"{prefix}/{k}": v.detach().numpy() for k, v in sd.items()}


def golden_scan():
    """``selective_scan_ref`` itself (``modules/mamba/selective_scan_interface.py:91-157``) in the
    reference's own [B, D, L] layout, with and without z / delta_bias / softplus."""
    ref = ref_shims.load_reference()
    g = torch.Generator().manual_seed(1234)
    B, D, L, N = 2, 24, 157, 16
    u = torch.randn(B, D, L, generator=g)
    delta = torch.randn(B, D, L, generator=g) * 0.5 - 3.0
    A = -torch.exp(torch.randn(D, N, generator=g) * 0.5 + 0.5)
    Bm = torch.randn(B, N, L, generator=g)
    Cm = torch.randn(B, N, L, generator=g)
    Dv = torch.randn(D, generator=g)
    z = torch.randn(B, D, L, generator=g)
    bias = torch.randn(D, generator=g) * 0.3
    out, last = ref.selective_scan_ref(u, delta, A, Bm, Cm, Dv, z, bias, True, True)
    out_nz, _ = ref.selective_scan_ref(u, delta, A, Bm, Cm, Dv, None, bias, True, True)
    # flipped-input call, exactly how bimamba.py:237 drives the backward direction
    out_b = ref.selective_scan_ref(u.flip(-1), delta.flip(-1), A, Bm.flip(-1), Cm.flip(-1), Dv,
                                   z.flip(-1), bias, True).flip(-1)
    np.savez_compressed(os.path.join(OUT, "scan_ref.npz"), u=u.numpy(), delta=delta.numpy(), A=A.numpy(),
                        B=Bm.numpy(), C=Cm.numpy(), D=Dv.numpy(), z=z.numpy(), bias=bias.numpy(),
                        out=out.numpy(), last_state=last.numpy(), out_nogate=out_nz.numpy(),
                        out_reverse=out_b.numpy())


def golden_forward(name, T, batch, seed, tag, own_init, bidirectional=True, mask_nonlinear="relu", rms_norm=True,
                   autocast=False):
    """``autocast``: run the reference's modules the way its recipes do by default (``precision: bf16``,
    hparams/WSJ0Mix/mambatasnet_S.yaml:38; ``torch.autocast`` region train_wsj0mix.py:161-164) -- on the CPU autocast
    backend here, whose op policy for this path (conv / linear / matmul in bf16, everything else in the dtype it is
    given) matches the CUDA one.  The outputs are stored as fp32 copies of the bf16 results."""
    import contextlib
    from dataclasses import replace
    hp = CONFIGS[name] if bidirectional else CONFIGS[name].causal()
    hp = replace(hp, rms_norm=rms_norm)
    ref = ref_shims.load_reference()
    enc, mask, dec = ref_shims.build_reference_model(hp.as_dict(), seed=seed, bidirectional=bidirectional,
                                                     mask_nonlinear=mask_nonlinear, rms_norm=rms_norm)
    if own_init:  # perturbed ("trained-like") weights, loaded strict into the reference modules
        sds = init_state_dicts(hp, seed)
        enc.load_state_dict(sds["encoder"], strict=True)
        mask.load_state_dict(sds["masknet"], strict=True)
        dec.load_state_dict(sds["decoder"], strict=True)
    mix, src = synth_mixture(batch, T, hp.sample_rate, seed=seed)
    taps = {}

    def hook(key):
        def fn(_m, _i, o):
            taps[key] = (o[0] if isinstance(o, tuple) else o).detach().float().numpy()
        return fn

    mask.mamba_net.layers[0].mixer.register_forward_hook(hook("mixer0_out"))
    mask.mamba_net.register_forward_hook(hook("stack_out"))
    mask.bottleneck_conv1x1.register_forward_hook(hook("bottleneck_out"))
    ctx = torch.autocast("cpu", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
    with torch.no_grad(), ctx:
        mix_w = enc(mix)
        est_mask = mask(mix_w)
        est = ref.compute_forward(enc, mask, dec, mix)
    if autocast:
        assert est.dtype == torch.bfloat16 and est_mask.dtype == torch.bfloat16
        taps = {k: v.astype(np.float32) if v.dtype != np.float32 else v for k, v in
                ((k, torch.as_tensor(v).float().numpy() if not isinstance(v, np.ndarray) else v) for k, v in taps.items())}
    f32 = lambda t: t.float().numpy()
    arrs = {"mix": mix.numpy(), "src": src.numpy(), "est": f32(est), "mix_w": f32(mix_w),
            "est_mask": f32(est_mask), "T": np.int64(T), "batch": np.int64(batch)}
    arrs.update({f"tap/{k}": v for k, v in taps.items()})
    arrs.update(_np(enc.state_dict(), "encoder"))
    arrs.update(_np(mask.state_dict(), "masknet"))
    arrs.update(_np(dec.state_dict(), "decoder"))
    np.savez_compressed(os.path.join(OUT, f"forward_{tag}.npz"), **arrs)
    print(tag, "est rms", float(est.pow(2).mean().sqrt()), "file",
          os.path.getsize(os.path.join(OUT, f"forward_{tag}.npz")) // 1024, "KiB")
    return enc, mask, dec


def golden_stream(mask, tag):
    """Streaming through the reference's own inference caches: ``MambaBlocksSequential.forward(x, inference_params)``
    (``modules/mamba_blocks.py:186-193``) -> prefill branch ``bimamba.py:271-304`` on the first ``L0`` tokens, then
    ``Mamba.step`` (``:320-372``) one token at a time.  Stored: input, the one-shot output, the streamed output and the
    final conv / ssm caches of every layer.  Weights are the ``masknet/mamba_net.*`` entries of forward_<tag>.npz."""
    ref = ref_shims.load_reference()
    net = mask.mamba_net
    g = torch.Generator().manual_seed(4321)
    B, L, L0 = 2, 41, 9
    h = torch.randn(B, L, net.norm_f.weight.shape[0], generator=g)
    with torch.no_grad():
        full = net(h)
        ip = ref.InferenceParams()
        outs = [net(h[:, :L0], inference_params=ip)]
        ip.seqlen_offset = L0
        for t in range(L0, L):
            outs.append(net(h[:, t:t + 1], inference_params=ip))
            ip.seqlen_offset += 1
        streamed = torch.cat(outs, dim=1)
    assert (streamed - full).abs().max() < 1e-5
    arrs = {"h": h.numpy(), "full": full.numpy(), "streamed": streamed.numpy(), "L0": np.int64(L0)}
    for i, (cs, ss) in ip.key_value_memory_dict.items():
        arrs[f"conv_state/{i}"] = cs.numpy()
        arrs[f"ssm_state/{i}"] = ss.numpy()
    np.savez_compressed(os.path.join(OUT, f"stream_{tag}.npz"), **arrs)
    print(f"stream_{tag}.npz", "max |streamed - one-shot|", float((streamed - full).abs().max()))


def golden_dp(tag, T, batch, seed, skip_around_intra, skip_n_block=0, n_dp=2):
    """DPMamba: the reference's ``Dual_Path_Model_Skip`` (vendored forward, ``modules/dual_path.py``) over the published
    speechbrain ``Dual_Path_Model`` members, with the real ``MambaBlocksSequential`` as intra / inter model."""
    from dataclasses import replace
    hp = replace(DP_CONFIGS["tiny"], skip_around_intra=skip_around_intra, skip_n_block=skip_n_block, n_dp=n_dp)
    ref = ref_shims.load_reference()
    enc, mask, dec = ref_shims.build_reference_dp_model(hp.as_dict(), seed=seed)
    sds = init_dp_state_dicts(hp, seed)
    enc.load_state_dict(sds["encoder"], strict=True)
    mask.load_state_dict(sds["masknet"], strict=True)
    dec.load_state_dict(sds["decoder"], strict=True)
    mix, src = synth_mixture(batch, T, hp.sample_rate, seed=seed)
    with torch.no_grad():
        mix_w = enc(mix)
        est_mask = mask(mix_w)
        est = ref.compute_forward(enc, mask, dec, mix)
    arrs = {"mix": mix.numpy(), "src": src.numpy(), "est": est.numpy(), "mix_w": mix_w.numpy(),
            "est_mask": est_mask.numpy(), "T": np.int64(T), "batch": np.int64(batch),
            "skip_around_intra": np.int64(skip_around_intra), "chunk_size": np.int64(hp.chunk_size),
            "n_dp": np.int64(hp.n_dp), "skip_n_block": np.int64(hp.skip_n_block)}
    arrs.update(_np(enc.state_dict(), "encoder"))
    arrs.update(_np(mask.state_dict(), "masknet"))
    arrs.update(_np(dec.state_dict(), "decoder"))
    np.savez_compressed(os.path.join(OUT, f"forward_{tag}.npz"), **arrs)
    print(tag, "est rms", float(est.pow(2).mean().sqrt()), "file",
          os.path.getsize(os.path.join(OUT, f"forward_{tag}.npz")) // 1024, "KiB")


def golden_si_snr():
    """``cal_si_snr`` of the reference itself (``baseline/avse2/utils/dnn.py:15-57``; plain torch, imported from the
    reference tree) on seeded (source, estimate) pairs: pins ``oracle.restate.cal_si_snr``."""
    import importlib.util
    path = os.path.join(ref_shims.REFERENCE_ROOT, "baseline", "avse2", "utils", "dnn.py")
    spec = importlib.util.spec_from_file_location("_ref_dnn", path)
    dnn = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(dnn)
    _, src = synth_mixture(3, 4001, seed=99)
    g = torch.Generator().manual_seed(5)
    est = src * 0.7 + torch.randn(src.shape, generator=g) * torch.tensor([0.002, 0.0002, 0.02]).view(3, 1, 1)
    est[0] = est[0].flip(-1)   # speakers swapped: channel-paired SI-SNR is very low, PIT must undo it
    est[0] += 0.3      # a DC offset must not matter (zero-mean)
    # reference layout is [T, B, C]; it modifies its estimate argument in place -> pass a clone
    neg = dnn.cal_si_snr(src.permute(1, 0, 2).contiguous(), est.permute(1, 0, 2).contiguous().clone())
    np.savez_compressed(os.path.join(OUT, "si_snr_ref.npz"), src=src.numpy(), est=est.numpy(),
                        si_snr=(-neg[0]).numpy())
    print("si_snr_ref.npz", (-neg[0]).tolist())


def golden_results_csv():
    """``test_results.csv`` exactly as ``Separation.save_results`` writes it (``Mamba-TasNet/train_wsj0mix.py:517-597``):
    the writer statements of the reference restated verbatim over fixed per-utterance metrics (the metrics themselves are
    pinned elsewhere), so that the product's writer can be compared byte for byte."""
    import csv
    g = np.random.default_rng(7)
    ids = ["440c0206_1.8_446o030c_-1.8", "22ga010b_0.5_050a050f_-0.5", "421a0108_2.1_01zo030f_-2.1"]
    sdr, sdr_i = g.normal(15, 3, 3), g.normal(15, 3, 3)
    sisnr, sisnr_i = g.normal(-14, 3, 3), g.normal(-14, 3, 3)      # the reference holds NEGATIVE si-snr (loss) here
    path = os.path.join(OUT, "test_results_ref.csv")
    all_sdrs, all_sdrs_i, all_sisnrs, all_sisnrs_i = [], [], [], []
    csv_columns = ["snt_id", "sdr", "sdr_i", "si-snr", "si-snr_i"]                      # :517
    with open(path, "w") as results_csv:                                                # :523
        writer = csv.DictWriter(results_csv, fieldnames=csv_columns)
        writer.writeheader()
        for i in range(3):
            row = {"snt_id": ids[i], "sdr": sdr[i], "sdr_i": sdr_i[i], "si-snr": -float(sisnr[i]),
                   "si-snr_i": -float(sisnr_i[i])}                                      # :577-583
            writer.writerow(row)
            all_sdrs.append(sdr[i])
            all_sdrs_i.append(sdr_i[i])
            all_sisnrs.append(-float(sisnr[i]))
            all_sisnrs_i.append(-float(sisnr_i[i]))
        row = {"snt_id": "avg", "sdr": np.array(all_sdrs).mean(), "sdr_i": np.array(all_sdrs_i).mean(),
               "si-snr": np.array(all_sisnrs).mean(), "si-snr_i": np.array(all_sisnrs_i).mean()}   # :591-597
        writer.writerow(row)
    np.savez(os.path.join(OUT, "test_results_ref_inputs.npz"), ids=np.array(ids), sdr=sdr, sdr_i=sdr_i,
             si_snr=-sisnr, si_snr_i=-sisnr_i)
    print("test_results_ref.csv", open(path).read().count("\n"), "lines")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    golden_scan()
    golden_forward("tiny", T=2000, batch=2, seed=1234, tag="tiny_refinit", own_init=False)
    golden_forward("tiny", T=1003, batch=3, seed=77, tag="tiny_trained", own_init=True)
    golden_si_snr()
    golden_results_csv()
    _, mask, _ = golden_forward("tiny", T=1203, batch=2, seed=55, tag="tiny_causal", own_init=True, bidirectional=False)
    golden_stream(mask, "tiny_causal")
    golden_forward("tiny", T=803, batch=1, seed=66, tag="tiny_softmax", own_init=True, mask_nonlinear="softmax")
    golden_forward("tiny", T=803, batch=1, seed=67, tag="tiny_layernorm", own_init=True, rms_norm=False)
    golden_forward("tiny", T=1003, batch=2, seed=78, tag="tiny_autocast", own_init=True, autocast=True)
    golden_dp("dp_tiny_skip", T=1203, batch=2, seed=91, skip_around_intra=True)
    golden_dp("dp_tiny_noskip", T=811, batch=1, seed=92, skip_around_intra=False)
    golden_dp("dp_tiny_blockskip", T=643, batch=1, seed=93, skip_around_intra=True, skip_n_block=1, n_dp=3)


if __name__ == "__main__":
    main()
