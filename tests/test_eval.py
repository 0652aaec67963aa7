"""CPU: the evaluation front end's host side -- wav I/O and the reference's ``test_results.csv`` format."""
import os
import wave

import numpy as np
import pytest
import torch

from avse_challenge_b200 import scoring, wavio


def test_results_csv_is_byte_identical_to_the_reference_writer(golden_dir, tmp_path):
    """``write_results_csv`` against ``tests/golden/test_results_ref.csv``, which oracle/make_golden.py wrote with the
    reference's own writer statements (``Mamba-TasNet/train_wsj0mix.py:517-597``)."""
    z = np.load(os.path.join(golden_dir, "test_results_ref_inputs.npz"))
    out = tmp_path / "test_results.csv"
    avg = scoring.write_results_csv(str(out), [str(i) for i in z["ids"]], z["si_snr"], z["si_snr_i"], z["sdr"], z["sdr_i"])
    assert out.read_bytes() == open(os.path.join(golden_dir, "test_results_ref.csv"), "rb").read()
    assert avg["snt_id"] == "avg" and abs(avg["si-snr"] - z["si_snr"].mean()) < 1e-12
    # without an SDR backend (mir_eval is not part of this path) the two SDR columns stay empty, the rest is unchanged
    scoring.write_results_csv(str(out), ["a", "b"], [1.0, 2.0], [3.0, 5.0])
    rows = out.read_text().splitlines()
    assert rows[0] == "snt_id,sdr,sdr_i,si-snr,si-snr_i" and rows[1] == "a,,,1.0,3.0" and rows[-1] == "avg,,,1.5,4.0"
    with pytest.raises(ValueError):
        scoring.write_results_csv(str(out), ["a"], [1.0, 2.0], [3.0])


@pytest.mark.parametrize("encoding,tol", [("float32", 0.0), ("pcm16", 1.0 / 32768)])
def test_wav_round_trip(tmp_path, encoding, tol):
    g = torch.Generator().manual_seed(1)
    for shape in ((1001,), (777, 2)):
        x = torch.randn(shape, generator=g) * 0.2
        p = str(tmp_path / f"x_{encoding}_{len(shape)}.wav")
        wavio.write_wav(p, x, 8000, encoding)
        y, rate = wavio.read_wav(p)
        assert rate == 8000 and y.shape == x.shape and y.dtype == torch.float32
        assert (x - y).abs().max().item() <= tol


def test_wav_reader_takes_stdlib_and_extensible_files(tmp_path):
    """Files written by another writer: the stdlib ``wave`` module (PCM16 / PCM32 / 24-bit) and a WAVE_FORMAT_EXTENSIBLE
    header with an odd-sized chunk in front of ``fmt`` (word alignment of chunks)."""
    import struct
    x = (np.sin(np.arange(400) * 0.1) * 20000).astype("<i2")
    p = str(tmp_path / "std16.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(x.tobytes())
    y, rate = wavio.read_wav(p)
    assert rate == 16000 and np.array_equal((y.numpy() * 32768).round().astype("<i2"), x)
    x24 = (np.sin(np.arange(100) * 0.3) * 8e6).astype(np.int32)
    p = str(tmp_path / "std24.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(1); w.setsampwidth(3); w.setframerate(8000)
        w.writeframes(b"".join(int(v).to_bytes(4, "little", signed=True)[:3] for v in x24))
    y, _ = wavio.read_wav(p)
    assert np.array_equal((y.numpy().astype(np.float64) * 8388608).round().astype(np.int32), x24)
    f = np.linspace(-0.5, 0.5, 64, dtype="<f4")
    fmt = struct.pack("<HHIIHH", 0xFFFE, 1, 8000, 32000, 4, 32) + struct.pack("<HHI", 22, 32, 4) + \
        struct.pack("<H", 3) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    body = b"LIST" + struct.pack("<I", 3) + b"abc\0" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + \
        b"data" + struct.pack("<I", f.nbytes) + f.tobytes()
    p = tmp_path / "ext.wav"
    p.write_bytes(b"RIFF" + struct.pack("<I", 4 + len(body)) + b"WAVE" + body)
    y, rate = wavio.read_wav(str(p))
    assert rate == 8000 and np.array_equal(y.numpy(), f)
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.wav"
        bad.write_bytes(b"RIFFxxxxWAVX")
        wavio.read_wav(str(bad))


def test_save_audio_layout(tmp_path):
    """``save_audio`` file names and peak normalisation (train_wsj0mix.py:606-642); needs no GPU."""
    g = torch.Generator().manual_seed(3)
    mix = torch.randn(1, 500, generator=g) * 0.1
    tgt, pred = torch.randn(1, 500, 2, generator=g) * 0.1, torch.randn(1, 500, 2, generator=g) * 0.3
    paths = scoring.save_audio(str(tmp_path), 7, mix, tgt, pred, 8000)
    names = sorted(os.path.basename(p) for p in paths)
    assert names == ["item7_mix.wav", "item7_source1.wav", "item7_source1hat.wav", "item7_source2.wav", "item7_source2hat.wav"]
    y, rate = wavio.read_wav(os.path.join(str(tmp_path), "audio_results", "item7_source2hat.wav"))
    assert rate == 8000 and abs(y.abs().max().item() - 1.0) < 1e-6
    assert torch.allclose(y, pred[0, :, 1] / pred[0, :, 1].abs().max())
