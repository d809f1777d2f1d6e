"""Golden fixtures: oracle drift guard + independent cross-check (CPU), CUDA parity (GPU)."""
import os

import numpy as np
import pytest

from oracle import slow_exact as sx
from parity import oracle_clip, compare_clip, assert_clip

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
KW = {
    "a": dict(n_fft=2048, hop_length=512, n_mels=128, n_mfcc=40, pad_mode="constant"),
    "b": dict(n_fft=512, hop_length=128, n_mels=40, n_mfcc=13, pad_mode="reflect"),
}


@pytest.mark.parametrize("case", ["a", "b"])
def test_oracle_reproduces_golden(case):
    y = G[f"{case}_y"]
    for i in range(len(y)):
        r = oracle_clip(y[i], **KW[case])
        assert np.array_equal(r["logmel"], G[f"{case}_logmel"][i])
        assert np.array_equal(r["mfcc"], G[f"{case}_mfcc"][i])
        assert np.array_equal(r["stats"], G[f"{case}_stats"][i])


def test_golden_matches_direct_dft():
    y = G["b_y"]
    for i in range(len(y)):
        ref = sx.features(y[i], 22050, 512, 128, 40, 13, "reflect")
        assert np.abs(G["b_logmel"][i] - ref["logmel"]).max() < 2e-3
        assert np.abs(G["b_mfcc"][i] - ref["mfcc"]).max() <= 1e-4 * np.abs(ref["mfcc"]).max()
        st = G["b_stats"][i]
        for j in (0, 1, 4):
            assert np.allclose(st[j], ref["stats"][j], rtol=2e-5, atol=1e-9)
        assert np.array_equal(st[3], ref["stats"][3])
        assert np.allclose(st[2], ref["stats"][2])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["a", "b"])
def test_cuda_matches_golden(case, built):
    import torch

    hl = built
    kw = KW[case]
    ex = hl.FeatureExtractor(ref=np.max, **kw)
    y = G[f"{case}_y"]
    out = ex.extract_device(torch.from_numpy(y).cuda())
    torch.cuda.synchronize()
    for i in range(len(y)):
        want = oracle_clip(y[i], **kw)   # for the rolloff tie rule only
        want.update(logmel=G[f"{case}_logmel"][i], mfcc=G[f"{case}_mfcc"][i], stats=G[f"{case}_stats"][i])
        got = {k: out[k][i].cpu().numpy() for k in ("logmel", "mfcc", "stats")}
        assert_clip(compare_clip(got, want, n_fft=kw["n_fft"], y=y[i]), where=f"golden {case}[{i}]")


# ---- fixtures from an independent code base (torchaudio's librosa-compatible transforms, float64) -----------
TA = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "torchaudio_v1.npz"))


def _ta_logmel(i):
    from oracle import librosa_oracle as orc

    return orc.power_to_db(TA["mel"][i].astype(np.float32), ref=np.max)


def test_oracle_matches_torchaudio_golden():
    from oracle import librosa_oracle as orc

    y = TA["y"]
    for i in range(len(y)):
        mel = orc.melspectrogram(y=y[i], sr=22050)
        assert np.abs(mel - TA["mel"][i]).max() <= 5e-6 * TA["mel"][i].max()      # float32 filterbank product
        mf = orc.mfcc(y=y[i], sr=22050, n_mfcc=40)
        assert np.abs(mf - TA["mfcc"][i]).max() <= 2e-5 * np.abs(TA["mfcc"][i]).max()
        cen = orc.spectral_centroid(y=y[i], sr=22050, pad_mode="reflect")[0]
        want = TA["centroid_reflect"][i]
        live = np.isfinite(want)                  # torchaudio: 0 / 0 on silent frames; librosa.util.normalize: 0
        assert np.all(cen[~live] == 0.0)
        assert np.abs(cen - want)[live].max() <= 1e-5 * np.abs(cen).max()


@pytest.mark.gpu
def test_cuda_matches_torchaudio_golden(built):
    """The CUDA path against a third party's numbers, no oracle in between: log-mel within the 0.01 dB budget
    (above the -80 dB floor, where a 1e-7 difference in the clip maximum cannot flip the clamp), MFCC and
    centroid within 1e-4 of the clip's largest value."""
    import torch

    hl = built
    y = TA["y"]
    out = hl.FeatureExtractor(ref=np.max, n_mfcc=40).extract_device(torch.from_numpy(y).cuda())
    refl = hl.FeatureExtractor(ref=np.max, n_mfcc=0, pad_mode="reflect").extract_device(torch.from_numpy(y).cuda())
    torch.cuda.synchronize()
    for i in range(len(y)):
        want = _ta_logmel(i)
        got = out["logmel"][i].cpu().numpy()
        assert got.shape == want.shape
        assert np.abs(got - want)[want > -79.9].max() <= 0.01
        assert np.abs(got - want).max() <= 0.011         # cells on the floor too: the clamp moves a cell by at most its own error
        mf = out["mfcc"][i].cpu().numpy()
        assert np.abs(mf - TA["mfcc"][i]).max() <= 1e-4 * np.abs(TA["mfcc"][i]).max()
        cen = refl["stats"][i, 0].cpu().numpy()
        want_c = TA["centroid_reflect"][i]
        live = np.isfinite(want_c)
        assert np.all(cen[~live] == 0.0)
        assert np.abs(cen - want_c)[live].max() <= 1e-4 * np.abs(want_c[live]).max()
