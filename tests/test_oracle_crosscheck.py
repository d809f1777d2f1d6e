"""Oracle vs an independent float64 direct-DFT implementation and vs torch / torchaudio; CPU only."""
import numpy as np
import pytest

from oracle import librosa_oracle as orc
from oracle import slow_exact as sx

SR = 22050


def _clip(seed, n=3000):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / SR
    y = 0.3 * np.sin(2 * np.pi * 440 * t) + 0.05 * rng.standard_normal(n)
    return y.astype(np.float32)


@pytest.mark.parametrize("mode", ["constant", "reflect", "edge"])
def test_oracle_matches_direct_dft(mode):
    y = _clip(1)
    kw = dict(n_fft=512, hop_length=128)
    ref = sx.features(y, SR, 512, 128, 40, 13, mode)
    S = np.abs(orc.stft(y, pad_mode=mode, **kw))
    assert np.abs(S - ref["S"]).max() <= 2e-5 * ref["S"].max()
    mel = orc.melspectrogram(y=y, sr=SR, n_mels=40, pad_mode=mode, **kw)
    assert np.abs(mel - ref["mel"]).max() <= 1e-5 * ref["mel"].max()
    lm = orc.power_to_db(mel, ref=np.max)
    assert np.abs(lm - ref["logmel"]).max() < 2e-3
    mf = orc.mfcc(y=y, sr=SR, n_mfcc=13, n_mels=40, pad_mode=mode, **kw)
    assert np.abs(mf - ref["mfcc"]).max() <= 1e-4 * np.abs(ref["mfcc"]).max()
    st = ref["stats"]
    assert np.allclose(orc.spectral_centroid(y=y, sr=SR, pad_mode=mode, **kw)[0], st[0], rtol=1e-5)
    assert np.allclose(orc.spectral_bandwidth(y=y, sr=SR, pad_mode=mode, **kw)[0], st[1], rtol=1e-5)
    assert np.allclose(orc.spectral_rolloff(y=y, sr=SR, pad_mode=mode, **kw)[0], st[2])
    assert np.array_equal(orc.zero_crossing_rate(y, frame_length=512, hop_length=128)[0], st[3])
    assert np.allclose(orc.rms(y=y, frame_length=512, hop_length=128, pad_mode=mode)[0], st[4], rtol=1e-5)


def test_filterbank_matches_triangle_definition():
    fb = orc.mel(sr=SR, n_fft=512, n_mels=40)
    assert np.abs(fb - sx.mel_filterbank(SR, 512, 40)).max() < 1e-7


def test_filterbank_matches_torchaudio():
    ta = pytest.importorskip("torchaudio")
    fb = orc.mel(sr=SR, n_fft=2048)
    fb2 = ta.functional.melscale_fbanks(1025, 0.0, SR / 2, 128, SR, norm="slaney", mel_scale="slaney").numpy().T
    assert np.abs(fb - fb2).max() < 5e-7


def test_stft_matches_torch():
    torch = pytest.importorskip("torch")
    y = _clip(2, 22050)
    D = orc.stft(y)
    Dt = torch.stft(torch.from_numpy(y).double(), 2048, 512, window=torch.hann_window(2048, dtype=torch.float64),
                    center=True, pad_mode="constant", return_complex=True).numpy()
    assert np.abs(D - Dt).max() <= 1e-5 * np.abs(Dt).max()


def _music(seed, n=22050):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / SR
    f0 = rng.uniform(110, 660)
    y = sum((0.4 / k) * np.sin(2 * np.pi * f0 * k * t + rng.uniform(0, 6)) for k in range(1, 9)) * np.exp(-1.5 * t)
    return (y + 0.01 * rng.standard_normal(n)).astype(np.float32)


def test_melspectrogram_matches_torchaudio():
    """torchaudio.transforms.MelSpectrogram with the slaney scale / norm is torchaudio's librosa-compatible
    configuration (float32 STFT): an implementation written by others against the same definition."""
    ta = pytest.importorskip("torchaudio")
    torch = pytest.importorskip("torch")
    tr = ta.transforms.MelSpectrogram(sample_rate=SR, n_fft=2048, hop_length=512, n_mels=128, power=2.0, center=True,
                                      pad_mode="constant", norm="slaney", mel_scale="slaney")
    for seed in (3, 4):
        y = _music(seed)
        want = tr(torch.from_numpy(y)).numpy()
        got = orc.melspectrogram(y=y, sr=SR)
        assert got.shape == want.shape == (128, 44)
        assert np.abs(got - want).max() <= 2e-5 * want.max()
        # the dB image the scripts store: well inside the 0.01 dB budget wherever it is not on the -80 dB floor
        a, b = orc.power_to_db(got, ref=np.max), orc.power_to_db(want, ref=np.max)
        assert np.abs(a - b)[b > -79.0].max() <= 5e-3


def test_mfcc_matches_torchaudio():
    """torchaudio.transforms.MFCC (log_mels=False) = DCT-II ortho of amplitude_to_DB(power mel, top_db=80): the
    chain librosa.feature.mfcc runs, from an independent code base."""
    ta = pytest.importorskip("torchaudio")
    torch = pytest.importorskip("torch")
    tr = ta.transforms.MFCC(sample_rate=SR, n_mfcc=40, dct_type=2, norm="ortho", log_mels=False,
                            melkwargs=dict(n_fft=2048, hop_length=512, n_mels=128, power=2.0, center=True,
                                           pad_mode="constant", norm="slaney", mel_scale="slaney"))
    for seed in (5, 6):
        y = _music(seed)
        want = tr(torch.from_numpy(y)).numpy()
        got = orc.mfcc(y=y, sr=SR, n_mfcc=40, n_fft=2048, hop_length=512)
        assert got.shape == want.shape == (40, 44)
        assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()


def test_spectral_centroid_matches_torchaudio():
    ta = pytest.importorskip("torchaudio")
    torch = pytest.importorskip("torch")
    y = _music(7)
    want = ta.functional.spectral_centroid(torch.from_numpy(y), SR, pad=0, window=torch.hann_window(2048), n_fft=2048,
                                           hop_length=512, win_length=2048).numpy()
    got = orc.spectral_centroid(y=y, sr=SR, pad_mode="reflect")[0]      # torchaudio's spectrogram pads by reflection
    assert got.shape == want.shape and np.abs(got - want).max() <= 1e-4 * np.abs(want).max()


@pytest.mark.parametrize("n_fft,hop", [(2048, 512), (1024, 256), (512, 100)])
def test_oracle_stft_matches_scipy_signal(n_fft, hop):
    """A third party's STFT: scipy.signal.stft with the periodic Hann window, zero 'boundary' extension of n_fft/2
    samples on both sides (librosa's center=True, pad_mode="constant") and no trailing padding frames the same
    samples at the same hops; its 'spectrum' scaling divides by sum(window)."""
    from scipy.signal import get_window, stft

    y = _clip(7, n=3 * n_fft + 123).astype(np.float64)
    win = get_window("hann", n_fft, fftbins=True)
    _f, _t, Z = stft(y, fs=SR, window=win, nperseg=n_fft, noverlap=n_fft - hop, nfft=n_fft, boundary="zeros",
                     padded=False, return_onesided=True, scaling="spectrum")
    D = orc.stft(y.astype(np.float32), n_fft=n_fft, hop_length=hop, pad_mode="constant")
    T = min(D.shape[1], Z.shape[1])           # scipy drops a trailing partial hop, librosa keeps 1 + n // hop frames
    assert T >= D.shape[1] - 1
    got, want = D[:, :T], Z[:, :T] * win.sum()
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
