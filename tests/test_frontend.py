"""librosa.load's arithmetic on device: int16 -> float32, mono mix, polyphase resampling,
crop / zero pad, per-clip lengths - against the oracle's restatement, whose resampler IS
scipy.signal.resample_poly (a third-party implementation: this row's parity is pinned)."""
import ctypes as C
import os
import wave

import numpy as np
import pytest

from oracle import librosa_oracle as orc

SR = 22050
WAVE_TOL = 2e-6          # absolute, samples in [-1, 1]: float32 accumulation order of ~44 taps


def test_resample_taps_are_scipys(built):
    """Host-side table: scipy.signal.resample_poly's default filter, bit for bit (no GPU needed)."""
    import scipy.signal as ss
    from hybrid_language_music_clustering_vae_b200 import _lib

    for sr_in, sr_out in ((44100, 22050), (48000, 22050), (16000, 22050), (8000, 22050), (32000, 22050), (22050, 16000)):
        buf = np.zeros(40000, np.float32)
        n = _lib.lib.hlmc_resample_taps(sr_in, sr_out, buf.ctypes.data, buf.size)
        g = np.gcd(sr_in, sr_out)
        up, down = sr_out // g, sr_in // g
        mr = max(up, down)
        ref = ss.firwin(20 * mr + 1, 1.0 / mr, window=("kaiser", 5.0)).astype(np.float32) * np.float32(up)
        assert n == len(ref)
        assert np.array_equal(buf[:n], ref)
        assert _lib.lib.hlmc_resampled_length(66150, sr_in, sr_out) == int(np.ceil(66150 * sr_out / sr_in))
    assert _lib.lib.hlmc_resample_taps(0, 22050, None, 0) < 0


def _pcm(n, ch, seed, sr):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    x = np.stack([0.3 * np.sin(2 * np.pi * (220.0 * (c + 1)) * t + c) + 0.05 * rng.standard_normal(n)
                  + 0.2 * np.sin(2 * np.pi * 9000.0 * t) for c in range(ch)], axis=1)
    return np.clip(np.rint(x * 32768.0), -32768, 32767).astype(np.int16)


@pytest.mark.gpu
@pytest.mark.parametrize("sr_in,ch", [(44100, 1), (44100, 2), (48000, 2), (16000, 1), (22050, 2), (8000, 3), (32000, 1)])
def test_frontend_matches_librosa_load_arithmetic(built, sr_in, ch):
    import torch

    hl = built
    ex = hl.FeatureExtractor(n_mfcc=0)
    n = int(0.7 * sr_in) + 13
    frames = np.stack([_pcm(n, ch, 10 + b, sr_in) for b in range(3)])
    want = np.stack([orc.load_pcm16(frames[b], sr_in, sr=SR)[0] for b in range(3)])
    got = ex.load_frontend_device(torch.from_numpy(frames).cuda(), sr_in=sr_in).cpu().numpy()
    assert got.shape == want.shape == (3, int(np.ceil(n * SR / sr_in)))
    assert np.abs(got - want).max() <= WAVE_TOL
    if sr_in == SR:                       # no filter: conversion and mono mix are exact
        assert np.array_equal(got, want)
    # float32 input takes the same path
    ff = frames.astype(np.float32) / np.float32(32768.0)
    got_f = ex.load_frontend_device(torch.from_numpy(ff).cuda(), sr_in=sr_in).cpu().numpy()
    assert np.abs(got_f - want).max() <= WAVE_TOL
    # padded, with per-clip lengths: each clip ends at ITS resampled length, then exact zeros
    valid = np.array([n, n // 2, 5], np.int64)
    pad_to = want.shape[1] + 100
    got_v = ex.load_frontend_device(torch.from_numpy(frames).cuda(), sr_in=sr_in, pad_to=pad_to,
                                    valid_frames=valid).cpu().numpy()
    for b in range(3):
        w = orc.load_pcm16(frames[b, :valid[b]], sr_in, sr=SR)[0]
        assert np.abs(got_v[b, :len(w)] - w).max() <= WAVE_TOL
        assert not got_v[b, len(w):].any()
    ex.close()


@pytest.mark.gpu
def test_host_pipeline_with_front_end_and_features(built):
    """44.1 kHz stereo PCM16 in host memory -> features, in ONE call; also returns the waveform it saw."""
    hl = built
    sr_in, n = 44100, 44100 * 2 + 7
    frames = np.stack([_pcm(n, 2, 40 + b, sr_in) for b in range(5)])
    valid = np.array([n, n, n - 1000, n // 3, n], np.int64)
    expected = SR * 3
    ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max)
    r = ex.extract_host(frames, sr_in=sr_in, valid_frames=valid, pad_to=expected, wave_out=True, pooled=True,
                        chunk_clips=2)
    assert r["wave"].shape == (5, expected)
    from parity import oracle_clip, compare_clip, assert_clip

    for b in range(5):
        y, _ = orc.load_pcm16(frames[b, :valid[b]], sr_in, sr=SR)
        y = np.pad(y, (0, expected - len(y)))
        assert np.abs(r["wave"][b] - y).max() <= WAVE_TOL
        # features of the waveform the device produced: the usual tolerances against the oracle on that waveform
        want = oracle_clip(r["wave"][b], n_mfcc=40)
        got = {k: r[k][b] for k in ("logmel", "mfcc", "stats")}
        assert_clip(compare_clip(got, want), where=f"clip {b}")
    h2d, _d2h = ex.last_transfer_bytes()
    assert h2d == 5 * n * 2 * 2                      # only the int16 frames crossed PCIe
    ex.close()


def _write_wav(path, frames, sr):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(frames.shape[1])
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.ascontiguousarray(frames, dtype="<i2").tobytes())


@pytest.mark.gpu
def test_process_single_file_and_load_audio_file(built, tmp_path):
    """[R] _advanced.py:158-183 / 1_preprocessing.py:137-153 on real WAV files (44.1 kHz stereo, 22.05 kHz mono,
    one longer than `duration`, one unreadable): same dict keys, shapes and failure records as the script."""
    hl = built
    pp = hl.preprocessing
    cfg = dict(pp.ADV_CONFIG, duration=2)                       # short clips keep the oracle fast
    files = []
    specs = [(44100, 2, 1.3), (22050, 1, 0.8), (48000, 1, 2.6), (44100, 2, 1.0)]
    for i, (sr, ch, secs) in enumerate(specs):
        fr = _pcm(int(secs * sr), ch, 70 + i, sr)
        p = tmp_path / f"{i}.wav"
        _write_wav(p, fr, sr)
        files.append({"path": str(p), "genre": "rock", "lyrics": "la", "language": "en", "filename": p.name,
                      "file_id": str(i), "_frames": fr, "_sr": sr})
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"not a wav file")
    files.append({"path": str(bad), "genre": "x", "lyrics": "", "language": "bn", "filename": "bad.wav", "file_id": "b"})
    res = pp.process_files_advanced(files, cfg)
    assert res[-1] == {"status": "failed", "path": str(bad), "error": "Load failed"}
    for info, r in zip(files[:-1], res[:-1]):
        assert r["status"] == "success"
        assert set(r) == {"status", "mel_spec", "flat_feat", "genre", "lyrics", "language", "filename", "file_id"}
        audio, sr = orc.load_audio_file_pcm16(info["_frames"], info["_sr"], cfg)
        assert sr == SR and len(audio) == SR * cfg["duration"]
        want_img = orc.adv_extract_mel_spectrogram(audio, sr, cfg)
        want_flat = orc.extract_flattened_features(audio, sr, cfg)
        assert r["mel_spec"].shape == (128, 1024) and r["mel_spec"].dtype == np.float32
        assert np.abs(r["mel_spec"] - want_img).max() <= 0.011     # 0.01 dB + what a 2e-6 input difference can move
        assert r["flat_feat"].shape == (290,) and r["flat_feat"].dtype == np.float64
        assert np.abs(r["flat_feat"][:256] - want_flat[:256]).max() <= 0.011
        assert np.allclose(r["flat_feat"][256:266], want_flat[256:266], rtol=2e-4, atol=SR / 2048 / 40)
        assert np.abs(r["flat_feat"][266:] - want_flat[266:]).max() <= 2e-4
    one = pp.process_single_file(files[0], cfg)
    assert one["status"] == "success" and np.array_equal(one["mel_spec"], res[0]["mel_spec"])
    a, sr = pp.load_audio_file(files[0]["path"], cfg)
    want, _ = orc.load_audio_file_pcm16(files[0]["_frames"], files[0]["_sr"], cfg)
    assert sr == SR and a.shape == want.shape and np.abs(a - want).max() <= WAVE_TOL
    assert pp.load_audio_file(str(bad), cfg) == (None, None)
