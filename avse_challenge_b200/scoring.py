"""Evaluation front end of the separator: per-utterance SI-SNR / SI-SNRi with PIT, on the device, and the
reference's ``test_results.csv`` format.

Mirrors ``Separation.save_results`` (``Mamba-TasNet/train_wsj0mix.py:503-604``): for every test utterance the
reference separates the mixture, evaluates ``-get_si_snr_with_pitwrapper(predictions, targets)`` [3P speechbrain; the
in-repo statement of the same SI-SNR is ``cal_si_snr``, ``baseline/avse2/utils/dnn.py:15-57``], the same for the
unprocessed mixture, and writes ``snt_id, sdr, sdr_i, si-snr, si-snr_i`` rows plus an ``avg`` row.  Here the SI-SNR
part runs as one streaming CUDA pass over (est, src, mix) per batch (``mtn_si_snr_pit_n_fwd``, 1..4 speakers); the SDR
columns of the reference come from ``mir_eval.separation.bss_eval_sources`` (a CPU library that is not part of this path
and not installed here): ``save_results`` fills them through ``sdr_fn`` when the caller has one and writes them empty
otherwise.  ``save_results`` / ``save_audio`` mirror the reference's evaluation loop (batch 1, one utterance after the
other, every length different) on top of any separator of this package; wav files go through ``wavio`` (no torchaudio).
"""
from __future__ import annotations

import csv
from typing import Iterable, Optional, Sequence

import torch

from . import _lib
from ._lib import check, ptr

CSV_COLUMNS = ["snt_id", "sdr", "sdr_i", "si-snr", "si-snr_i"]          # train_wsj0mix.py:517


def si_snr_pit(est: torch.Tensor, src: torch.Tensor, mix: torch.Tensor) -> dict:
    """``est``, ``src`` [B, T, n] fp32 CUDA (n = 1..4 speakers; the recipes use 2 or 3, ``mambatasnet_S.yaml:39``), ``mix``
    [B, T] fp32 CUDA -> dict of [B] tensors (dB, larger = better): ``si_snr`` (best of the n! assignments, mean over the
    speakers), ``si_snr_i`` (minus the mixture's SI-SNR), ``perm`` (lexicographic rank of the best assignment; n = 2: 0 =
    est0<->src0, 1 = est0<->src1), ``baseline``, ``pairs`` [B, n, n] (est i vs src j) and ``assignment`` [B, n] (the
    source matched to estimate i)."""
    for t in (est, src, mix):
        if not t.is_cuda or t.dtype != torch.float32:
            raise _lib.MtnError("si_snr_pit expects fp32 CUDA tensors (no CPU fallback)")
    if est.dim() != 3 or not 1 <= est.shape[-1] <= 4 or src.shape != est.shape or mix.shape != est.shape[:2]:
        raise _lib.MtnError(f"si_snr_pit: shapes est {tuple(est.shape)} src {tuple(src.shape)} mix {tuple(mix.shape)}; "
                            "need [B, T, n], [B, T, n], [B, T] with n = 1..4")
    est, src = est.contiguous(), src.contiguous()
    if mix.stride(1) != 1:
        mix = mix.contiguous()
    B, T = mix.shape
    n = est.shape[-1]
    lib = _lib.load()
    nbytes = int(lib.mtn_si_snr_workspace_bytes_n(B, T, n))
    work = torch.empty(nbytes // 8, dtype=torch.float64, device=est.device)
    stride = 4 + n * n + n
    out = torch.empty((B, stride), dtype=torch.float32, device=est.device)
    check(lib.mtn_si_snr_pit_n_fwd(ptr(est), ptr(src), ptr(mix), mix.stride(0), B, T, n, ptr(work), nbytes, ptr(out), stride,
                                   torch.cuda.current_stream().cuda_stream), "mtn_si_snr_pit_n_fwd")
    return {"si_snr": out[:, 0], "si_snr_i": out[:, 1], "perm": out[:, 2].to(torch.int64), "baseline": out[:, 3],
            "pairs": out[:, 4:4 + n * n].reshape(B, n, n), "assignment": out[:, 4 + n * n:].to(torch.int64)}


def write_results_csv(path: str, snt_ids: Sequence[str], si_snr: Iterable[float], si_snr_i: Iterable[float],
                      sdr: Optional[Iterable[float]] = None, sdr_i: Optional[Iterable[float]] = None) -> dict:
    """Write the reference's ``test_results.csv`` (``train_wsj0mix.py:517-597``): one row per utterance and a final
    ``avg`` row.  Returns the averages."""
    si_snr, si_snr_i = [float(v) for v in si_snr], [float(v) for v in si_snr_i]
    sdr = [float(v) for v in sdr] if sdr is not None else None
    sdr_i = [float(v) for v in sdr_i] if sdr_i is not None else None
    n = len(snt_ids)
    if not (len(si_snr) == len(si_snr_i) == n) or (sdr is not None and len(sdr) != n) or (sdr_i is not None and len(sdr_i) != n):
        raise ValueError("write_results_csv: column lengths differ")
    import numpy as np
    mean = lambda v: np.array(v).mean() if v else float("nan")          # :591-596 (np.array(all_x).mean())
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=CSV_COLUMNS)
        w.writeheader()
        for i, sid in enumerate(snt_ids):
            w.writerow({"snt_id": sid, "sdr": "" if sdr is None else sdr[i], "sdr_i": "" if sdr_i is None else sdr_i[i],
                        "si-snr": si_snr[i], "si-snr_i": si_snr_i[i]})
        avg = {"snt_id": "avg", "sdr": "" if sdr is None else mean(sdr), "sdr_i": "" if sdr_i is None else mean(sdr_i),
               "si-snr": mean(si_snr), "si-snr_i": mean(si_snr_i)}
        w.writerow(avg)
    return avg


def save_audio(save_folder: str, snt_id, mixture: torch.Tensor, targets: torch.Tensor, predictions: torch.Tensor,
               sample_rate: int) -> list:
    """``Separation.save_audio`` (``train_wsj0mix.py:606-642``): per speaker the estimate and the target, and the mixture,
    each scaled to peak 1, as ``<save_folder>/audio_results/item{snt_id}_source{k}hat.wav`` / ``_source{k}.wav`` /
    ``_mix.wav``.  ``mixture`` [1, T], ``targets`` / ``predictions`` [1, T, n_spk].  Returns the paths written."""
    import os
    from .wavio import write_wav
    save_path = os.path.join(save_folder, "audio_results")
    os.makedirs(save_path, exist_ok=True)
    paths = []

    def put(name, signal):
        signal = signal / signal.abs().max()                      # :618, :629, :639 (no epsilon in the reference either)
        p = os.path.join(save_path, name)
        write_wav(p, signal, sample_rate)
        paths.append(p)

    for ns in range(predictions.shape[-1]):
        put(f"item{snt_id}_source{ns + 1}hat.wav", predictions[0, :, ns])
        put(f"item{snt_id}_source{ns + 1}.wav", targets[0, :, ns])
    put(f"item{snt_id}_mix.wav", mixture[0])
    return paths


def save_results(separator, test_items, output_folder: str, num_spks: int = 2, device="cuda", sdr_fn=None,
                 n_audio_files: int = 0, sample_rate: int = 8000) -> dict:
    """``Separation.save_results`` (``train_wsj0mix.py:503-604``): separate every test utterance (batch 1, as the reference's
    loader does), score it and write ``<output_folder>/test_results.csv`` (``snt_id, sdr, sdr_i, si-snr, si-snr_i`` rows plus
    the ``avg`` row).  Returns the averages.

    ``separator``: any ``[1, T] -> [1, T, num_spks]`` callable of this package (``MambaTasNetSeparator``, ``DPMambaSeparator``,
    ``SeparatorEngine`` ...).  Build it with ``use_graph=False`` for a test set whose utterances all differ in length: a CUDA
    graph per length would cost an eager run plus a capture each (the engines bound their per-shape caches, LRU).
    ``test_items``: iterable of dicts ``{"id", "mix": tensor [T] or wav path, "sources": [tensor or path] * num_spks}``
    (the ``mix_sig`` / ``s1_sig`` / ``s2_sig`` (/ ``s3_sig``) fields of the reference's data pipeline, ``:680-713``).
    ``sdr_fn(targets [n, T] numpy, estimates [n, T] numpy) -> per-source SDR array``: e.g. ``lambda r, e:
    mir_eval.separation.bss_eval_sources(r, e)[0]`` where mir_eval exists (``:564-572``); ``None`` leaves the SDR columns
    empty.  ``n_audio_files``: also ``save_audio`` the first n utterances (``:478-485`` ``n_audio_to_save``)."""
    import os
    from .wavio import read_wav
    os.makedirs(output_folder, exist_ok=True)

    def load(x):
        if isinstance(x, str):
            sig, rate = read_wav(x)
            if rate != sample_rate:
                raise ValueError(f"{x}: sample rate {rate} != {sample_rate} (the reference resamples offline)")
            x = sig if sig.dim() == 1 else sig[:, 0]
        return x.to(device=device, dtype=torch.float32)

    ids, sisnr, sisnr_i, sdr, sdr_i = [], [], [], [], []
    for n_done, item in enumerate(test_items):
        mix = load(item["mix"]).unsqueeze(0)                                          # [1, T]
        targets = torch.stack([load(s) for s in item["sources"][:num_spks]], dim=-1).unsqueeze(0)   # :540-543
        with torch.no_grad():
            predictions = separator(mix)                                              # :546
        if predictions.shape != targets.shape:
            raise ValueError(f"separator returned {tuple(predictions.shape)}, targets are {tuple(targets.shape)}")
        sc = si_snr_pit(predictions, targets, mix)                                    # :548-558 in one device pass
        ids.append(str(item["id"]))
        sisnr.append(sc["si_snr"].item())
        sisnr_i.append(sc["si_snr_i"].item())
        if sdr_fn is not None:                                                        # :564-574
            ref_np = targets[0].t().cpu().numpy()
            s = float(sdr_fn(ref_np, predictions[0].t().cpu().numpy()).mean())
            base = float(sdr_fn(ref_np, torch.stack([mix[0]] * num_spks).cpu().numpy()).mean())
            sdr.append(s)
            sdr_i.append(s - base)
        if n_done < n_audio_files:
            save_audio(output_folder, ids[-1], mix, targets, predictions, sample_rate)
    return write_results_csv(os.path.join(output_folder, "test_results.csv"), ids, sisnr, sisnr_i,
                             sdr if sdr_fn is not None else None, sdr_i if sdr_fn is not None else None)
