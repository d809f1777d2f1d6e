"""Known-answer tests that pin the oracle; CPU only."""
import numpy as np
import pytest

from oracle import librosa_oracle as orc

SR = 22050


@pytest.mark.parametrize("n,T", [(66150, 130), (661500, 1292), (2047, 4), (2048, 5), (2049, 5), (511, 1)])
def test_frame_counts(n, T):
    y = np.zeros(n, np.float32)
    assert orc.num_frames(n) == T
    assert orc.stft(y).shape == (1025, T)
    assert orc.zero_crossing_rate(y).shape == (1, T)
    assert orc.rms(y=y).shape == (1, T)
    assert orc.spectral_rolloff(y=y, sr=SR).shape == (1, T)


def test_all_zero_clip():
    y = np.zeros(22050, np.float32)
    lm = orc.basic_extract_mel_spectrogram(y, SR)
    assert lm.dtype == np.float32 and np.all(lm == 0.0)
    mf = orc.basic_extract_mfcc(y, SR)
    assert np.allclose(mf[0], -100.0 * np.sqrt(128), rtol=1e-6)
    assert np.abs(mf[1:]).max() < 1e-3
    sp = orc.extract_spectral_features(y, SR)
    for k, v in sp.items():
        assert np.all(v == 0.0), k


def test_on_bin_cosine():
    N, k0, A = 2048, 64, 0.5
    t = np.arange(66150)
    y = (A * np.cos(2 * np.pi * k0 * t / N)).astype(np.float32)
    D = np.abs(orc.stft(y))[:, 50]
    assert abs(D[k0] - A * N / 4) < 1e-3 and abs(D[k0 - 1] - A * N / 8) < 1e-3 and abs(D[k0 + 1] - A * N / 8) < 1e-3
    rest = np.delete(D, [k0 - 1, k0, k0 + 1])
    assert rest.max() < 1e-4
    cen = orc.spectral_centroid(y=y, sr=SR)[0, 50]
    assert abs(cen - k0 * SR / N) / (k0 * SR / N) < 1e-6
    # cumulative shares 0.25 / 0.75 / 1.0 -> first >= 0.85 is bin k0+1
    assert abs(orc.spectral_rolloff(y=y, sr=SR)[0, 50] - (k0 + 1) * SR / N) < 1e-6
    assert abs(orc.rms(y=y)[0, 50] - A / np.sqrt(2)) < 1e-6


def test_constant_and_alternating():
    c = 0.25
    y = np.full(22050, c, np.float32)
    D = np.abs(orc.stft(y))[:, 20]
    assert abs(D[0] - c * 1024) < 1e-3 and abs(D[1] - c * 512) < 1e-3 and D[2:].max() < 1e-4
    assert np.all(orc.zero_crossing_rate(y) == 0.0)
    assert np.allclose(orc.rms(y=y)[0, 5:-5], c)
    alt = np.where(np.arange(22050) % 2 == 0, 1.0, -1.0).astype(np.float32)
    assert orc.zero_crossing_rate(alt)[0, 20] == 2047 / 2048


def test_impulse_is_flat():
    y = np.zeros(22050, np.float32)
    s = 5000
    y[s] = 1.0
    t = 10
    D = np.abs(orc.stft(y))[:, t]
    w = 0.5 - 0.5 * np.cos(2 * np.pi * (s - t * 512 + 1024) / 2048)
    assert np.allclose(D, w, atol=1e-6)


def test_refmax_range_and_floor():
    rng = np.random.default_rng(3)
    y = (rng.standard_normal(22050) * np.logspace(0, -7, 22050)).astype(np.float32)
    lm = orc.basic_extract_mel_spectrogram(y, SR)
    assert lm.max() == 0.0 and lm.min() >= -80.0
    assert np.any(lm == -80.0)


def test_filterbank_and_dct_properties():
    fb = orc.mel(sr=SR, n_fft=2048)
    assert fb.shape == (128, 1025) and fb.dtype == np.float32
    assert (fb > 0).sum() == 2018 and np.all(fb[:, 0] == 0) and np.all(fb[:, 1024] == 0)
    assert abs(fb.max() - 0.038421705) < 1e-8
    wide = fb[100:].sum(axis=1) * (SR / 2048)
    assert np.allclose(wide, 1.0, atol=0.02)
    x = np.random.default_rng(0).standard_normal((128, 7)).astype(np.float32)
    import scipy.fftpack
    full = scipy.fftpack.dct(x, axis=0, type=2, norm="ortho")
    assert np.allclose(np.linalg.norm(full, axis=0), np.linalg.norm(x, axis=0), rtol=1e-5)
    assert np.allclose(full[0], x.sum(axis=0) / np.sqrt(128), rtol=1e-5)


def test_pad_modes_only_touch_edge_frames():
    y = (np.random.default_rng(5).standard_normal(66150) * 0.1).astype(np.float32)
    a = np.abs(orc.stft(y, pad_mode="constant"))
    b = np.abs(orc.stft(y, pad_mode="reflect"))
    assert np.array_equal(a[:, 2:-2], b[:, 2:-2])
    for t in (0, 1, 128, 129):
        assert not np.array_equal(a[:, t], b[:, t])


def test_mfcc_fusion_identity():
    """mfcc == DCT(logmel_refmax) with a constant added to coefficient 0."""
    import scipy.fftpack
    y = (np.random.default_rng(6).standard_normal(22050) * 0.1).astype(np.float32)
    mel = orc.melspectrogram(y=y, sr=SR)
    ref = orc.mfcc(y=y, sr=SR, n_mfcc=40)
    alt = scipy.fftpack.dct(orc.power_to_db(mel, ref=np.max), axis=-2, type=2, norm="ortho")[:40]
    alt[0] += 10 * np.log10(max(1e-10, mel.max())) * np.sqrt(128)
    assert np.abs(alt - ref).max() < 1e-3


def test_script_shapes():
    y = (np.random.default_rng(7).standard_normal(66150) * 0.1).astype(np.float32)
    assert orc.extract_all_features(y, SR).shape == (370,)
    assert orc.extract_all_features(y, SR).dtype == np.float64
    assert orc.extract_flattened_features(y, SR).shape == (290,)
    m = orc.adv_extract_mel_spectrogram(y, SR)
    assert m.shape == (128, 1024) and m.dtype == np.float32
    assert np.all(m[:, 130:] == m[:, :130].min())
    long = np.zeros(661500, np.float32)
    long[:66150] = y
    assert orc.adv_extract_mel_spectrogram(long, SR).shape == (128, 1024)


def test_parameter_errors():
    y = np.zeros(4096, np.float32)
    with pytest.raises(orc.ParameterError):
        orc.power_to_db(np.ones((4, 4), np.float32), amin=0)
    with pytest.raises(orc.ParameterError):
        orc.power_to_db(np.ones((4, 4), np.float32), top_db=-1)
    with pytest.raises(orc.ParameterError):
        orc.spectral_rolloff(y=y, roll_percent=1.0)
    with pytest.raises(orc.ParameterError):
        orc.stft(np.array([0.0, np.nan], np.float32))
    with pytest.raises(orc.ParameterError):
        orc.stft(y, center=False, n_fft=8192)


def test_real_librosa_if_present():
    librosa = pytest.importorskip("librosa")
    y = (np.random.default_rng(8).standard_normal(22050) * 0.1).astype(np.float32)
    a = librosa.power_to_db(librosa.feature.melspectrogram(y=y, sr=SR), ref=np.max)
    assert np.abs(a - orc.basic_extract_mel_spectrogram(y, SR)).max() < 1e-3
    assert np.abs(librosa.feature.mfcc(y=y, sr=SR, n_mfcc=40) - orc.mfcc(y=y, sr=SR, n_mfcc=40)).max() < 1e-2
