"""RIFF / WAVE reader and writer without third-party packages (the reference goes through torchaudio:
``sb.dataio.dataio.read_audio`` in ``Mamba-TasNet/train_wsj0mix.py:680-713`` and ``torchaudio.save`` in ``save_audio``,
``:606-642``).

Formats: PCM 16 / 24 / 32 bit (``wFormatTag`` 1) and IEEE float 32 (``wFormatTag`` 3, what ``torchaudio.save`` writes for a
float tensor), either plain or ``WAVE_FORMAT_EXTENSIBLE`` headers; any channel count.  Samples come back as float32 in
[-1, 1) -- ``[T]`` for mono (what ``read_audio`` returns for the wsj0-mix files), ``[T, C]`` otherwise.
"""
from __future__ import annotations

import struct
from typing import Tuple

import numpy as np
import torch

_PCM, _FLOAT, _EXT = 1, 3, 0xFFFE


def read_wav(path: str) -> Tuple[torch.Tensor, int]:
    """``(samples float32 [T] or [T, C], sample_rate)``."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, rate, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == _EXT and len(body) >= 26:
                tag = struct.unpack("<H", body[24:26])[0]          # first two bytes of the SubFormat GUID
            fmt = (tag, ch, rate, bits)
        elif cid == b"data":
            payload = body
        pos += 8 + size + (size & 1)                                # chunks are word aligned
    if fmt is None or payload is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, ch, rate, bits = fmt
    if tag == _FLOAT and bits == 32:
        x = np.frombuffer(payload, dtype="<f4").astype(np.float32)
    elif tag == _PCM and bits == 16:
        x = np.frombuffer(payload, dtype="<i2").astype(np.float32) / 32768.0
    elif tag == _PCM and bits == 32:
        x = (np.frombuffer(payload, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    elif tag == _PCM and bits == 24:
        b = np.frombuffer(payload[: len(payload) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        x = (v / 8388608.0).astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported wav format tag {tag} with {bits} bits")
    x = x[: len(x) // ch * ch]
    t = torch.from_numpy(x.copy())
    return (t if ch == 1 else t.view(-1, ch)), int(rate)


def write_wav(path: str, samples: torch.Tensor, sample_rate: int, encoding: str = "float32") -> None:
    """``samples`` float ``[T]`` or ``[T, C]`` (CPU or CUDA).  ``encoding``: ``"float32"`` (IEEE float, torchaudio's default
    for float tensors) or ``"pcm16"`` (clipped to [-1, 1), like torchaudio's ``encoding="PCM_S", bits_per_sample=16``)."""
    x = samples.detach().to("cpu", torch.float32)
    if x.dim() == 1:
        x = x.unsqueeze(1)
    if x.dim() != 2:
        raise ValueError("write_wav expects [T] or [T, channels]")
    T, ch = x.shape
    if encoding == "float32":
        tag, bits, payload = _FLOAT, 32, x.contiguous().numpy().astype("<f4").tobytes()
    elif encoding == "pcm16":
        q = torch.clamp(torch.round(x * 32768.0), -32768, 32767).to(torch.int16)
        tag, bits, payload = _PCM, 16, q.contiguous().numpy().astype("<i2").tobytes()
    else:
        raise ValueError("encoding must be 'float32' or 'pcm16'")
    block = ch * bits // 8
    fmt = struct.pack("<HHIIHH", tag, ch, sample_rate, sample_rate * block, block, bits)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt
    if tag == _FLOAT:                                               # non-PCM formats carry a fact chunk
        chunks += b"fact" + struct.pack("<II", 4, T)
    chunks += b"data" + struct.pack("<I", len(payload)) + payload + (b"\0" if len(payload) & 1 else b"")
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks)
