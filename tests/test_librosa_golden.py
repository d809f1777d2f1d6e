"""Activates when tests/golden/librosa_v1.npz exists (made by tests/golden/make_librosa_golden.py on a machine
that has real librosa): the oracle (CPU) and the CUDA path (GPU) against librosa's own numbers, at north_star's
tolerances.  This is the route from "parity unpinned" to a pinned oracle; without the file the tests skip."""
import os

import numpy as np
import pytest

from oracle import librosa_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden", "librosa_v1.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="no librosa fixture committed (librosa is not installable here)")
SR = 22050
KW = {
    "a": dict(n_fft=2048, hop_length=512, n_mels=128, n_mfcc=40, pad_mode="constant"),
    "b": dict(n_fft=512, hop_length=128, n_mels=40, n_mfcc=13, pad_mode="reflect"),
    "t": dict(n_fft=2048, hop_length=512, n_mels=128, n_mfcc=40, pad_mode="constant"),
}


def _inputs():
    g = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
    t = np.load(os.path.join(HERE, "golden", "torchaudio_v1.npz"))
    return {"a": g["a_y"], "b": g["b_y"], "t": t["y"]}


def _check(case, i, L, got):
    """got: dict with logmel, mfcc, stats (5, T) [, chroma, tuning] for clip i of `case`."""
    kw = KW[case]
    assert np.abs(got["logmel"] - L[f"{case}_logmel"][i]).max() <= 0.01
    mf = L[f"{case}_mfcc"][i]
    assert np.abs(got["mfcc"] - mf).max() <= 1e-4 * max(np.abs(mf).max(), 1e-6)
    for j, name in enumerate(("centroid", "bandwidth", "rolloff", "zcr", "rms")):
        want = L[f"{case}_{name}"][i][0]
        if name == "rolloff":
            assert (np.abs(got["stats"][j] - want) > 1e-3).sum() <= max(1, len(want) // 200)
        else:
            assert np.abs(got["stats"][j] - want).max() <= 1e-4 * max(np.abs(want).max(), 1e-12), name
    if "chroma" in got and f"{case}_chroma" in L:
        if abs(float(got["tuning"]) - float(L[f"{case}_tuning"][i])) < 1e-6:
            assert np.abs(got["chroma"] - L[f"{case}_chroma"][i]).max() <= 1e-4


@pytest.mark.parametrize("case", ["a", "b", "t"])
def test_oracle_matches_real_librosa(case):
    from parity import oracle_clip

    L = np.load(PATH)
    ys = _inputs()[case]
    for i, y in enumerate(ys):
        o = oracle_clip(y, **KW[case])
        got = {"logmel": o["logmel"], "mfcc": o["mfcc"], "stats": o["stats"]}
        if KW[case]["n_fft"] == 2048 and KW[case]["pad_mode"] == "constant":
            got["tuning"] = orc.estimate_tuning(S=np.abs(orc.stft(y)) ** 2, sr=SR, n_fft=2048)
            got["chroma"] = orc.chroma_stft(y=y, sr=SR)
            assert np.allclose(orc.extract_all_features(y, SR), L[f"{case}_all370"][i], rtol=1e-4, atol=1e-3)
        assert np.abs(orc.stft(y, n_fft=KW[case]["n_fft"], hop_length=KW[case]["hop_length"],
                               pad_mode=KW[case]["pad_mode"]) - L[f"{case}_stft"][i]).max() <= 1e-4
        _check(case, i, L, got)
    for sr_in in (44100, 48000, 16000):
        x = L[f"resample_in_{sr_in}"]
        assert np.abs(orc.resample(x, orig_sr=sr_in, target_sr=SR) - L[f"resample_polyphase_{sr_in}"]).max() <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["a", "b", "t"])
def test_cuda_matches_real_librosa(built, case):
    import torch

    L = np.load(PATH)
    ys = _inputs()[case]
    kw = KW[case]
    ex = built.FeatureExtractor(ref=np.max, **kw)
    chroma = kw["n_fft"] == 2048 and kw["pad_mode"] == "constant"
    out = ex.extract_device(torch.from_numpy(np.ascontiguousarray(ys)).cuda(), chroma=chroma)
    for i in range(len(ys)):
        got = {k: out[k][i].cpu().numpy() for k in ("logmel", "mfcc", "stats")}
        if chroma:
            got["chroma"], got["tuning"] = out["chroma"][i].cpu().numpy(), out["tuning"][i].cpu().numpy()
        _check(case, i, L, got)
    for sr_in in (44100, 48000, 16000):
        x = torch.from_numpy(L[f"resample_in_{sr_in}"]).cuda()[None]
        w = ex.load_frontend_device(x, sr_in=sr_in).cpu().numpy()[0]
        assert np.abs(w - L[f"resample_polyphase_{sr_in}"]).max() <= 2e-6
