"""Evaluation front end of the separator: per-utterance SI-SNR / SI-SNRi with PIT, on the device, and the
reference's ``test_results.csv`` format.

Mirrors ``Separation.save_results`` (``Mamba-TasNet/train_wsj0mix.py:503-604``): for every test utterance the
reference separates the mixture, evaluates ``-get_si_snr_with_pitwrapper(predictions, targets)`` [3P speechbrain; the
in-repo statement of the same SI-SNR is ``cal_si_snr``, ``baseline/avse2/utils/dnn.py:15-57``], the same for the
unprocessed mixture, and writes ``snt_id, sdr, sdr_i, si-snr, si-snr_i`` rows plus an ``avg`` row.  Here the SI-SNR
part runs as one streaming CUDA pass over (est, src, mix) per batch (``mtn_si_snr_pit_fwd``); the SDR columns of the
reference come from ``mir_eval.separation.bss_eval_sources`` (a CPU library that is not part of this path) and are
written empty unless the caller supplies them.
"""
from __future__ import annotations

import csv
from typing import Iterable, Optional, Sequence

import torch

from . import _lib
from ._lib import check, ptr

CSV_COLUMNS = ["snt_id", "sdr", "sdr_i", "si-snr", "si-snr_i"]          # train_wsj0mix.py:517


def si_snr_pit(est: torch.Tensor, src: torch.Tensor, mix: torch.Tensor) -> dict:
    """``est``, ``src`` [B, T, 2] fp32 CUDA, ``mix`` [B, T] fp32 CUDA -> dict of [B] tensors (dB, larger = better):
    ``si_snr`` (best permutation, mean over the two speakers), ``si_snr_i`` (minus the mixture's SI-SNR),
    ``perm`` (0: est0<->src0, 1: est0<->src1), ``baseline``, and ``pairs`` [B, 2, 2] (est i vs src j)."""
    for t in (est, src, mix):
        if not t.is_cuda or t.dtype != torch.float32:
            raise _lib.MtnError("si_snr_pit expects fp32 CUDA tensors (no CPU fallback)")
    if est.dim() != 3 or est.shape[-1] != 2 or src.shape != est.shape or mix.shape != est.shape[:2]:
        raise _lib.MtnError(f"si_snr_pit: shapes est {tuple(est.shape)} src {tuple(src.shape)} mix {tuple(mix.shape)}; "
                            "need [B, T, 2], [B, T, 2], [B, T]")
    est, src = est.contiguous(), src.contiguous()
    if mix.stride(1) != 1:
        mix = mix.contiguous()
    B, T = mix.shape
    lib = _lib.load()
    nbytes = int(lib.mtn_si_snr_workspace_bytes(B, T))
    work = torch.empty(nbytes // 8, dtype=torch.float64, device=est.device)
    out = torch.empty((B, 8), dtype=torch.float32, device=est.device)
    check(lib.mtn_si_snr_pit_fwd(ptr(est), ptr(src), ptr(mix), mix.stride(0), B, T, ptr(work), nbytes, ptr(out),
                                 torch.cuda.current_stream().cuda_stream), "mtn_si_snr_pit_fwd")
    return {"si_snr": out[:, 0], "si_snr_i": out[:, 1], "perm": out[:, 2].to(torch.int64), "baseline": out[:, 3],
            "pairs": out[:, 4:8].reshape(B, 2, 2)}


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
    mean = lambda v: (sum(v) / len(v)) if v else float("nan")
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
