"""CPU oracle: a numpy/scipy restatement of the librosa functions on the hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may.  It exists to answer "what would the reference have produced?".

PARITY UNPINNED: the arithmetic of the reference's hot path lives in librosa,
an un-vendored third-party dependency with no version pin ([R]
src/1_preprocessing.py:7, src/1_preprocessing_advanced.py:7 ``import librosa``;
no requirements file exists).  librosa is not installable in this image (no
network, no wheel on disk) and the reference holds no tests or golden vectors
for this path.  This file therefore restates the
published algorithm of librosa >= 0.10 (the API generation the scripts target:
keyword-only ``y=``/``sr=``, ``pad_mode="constant"`` default), built on the
same primitives librosa itself delegates to -- ``scipy.signal.get_window``,
``scipy.fft.rfft``, ``scipy.fftpack.dct`` -- following its operation order and
dtypes.  It is pinned only by (a) the analytic known-answer tests in
``tests/test_oracle_kat.py``, (b) an independent float64 direct-DFT
implementation (``oracle/slow_exact.py``), (c) a torchaudio filterbank /
torch.stft cross-check where torchaudio is importable, and (d) a
``pytest.importorskip("librosa")`` test that activates if a real librosa is
ever present.

Call sites in the reference that each function stands in for:

=====================  ========================================================
oracle function         reference call site
=====================  ========================================================
stft                    implicit in every feature call below
melspectrogram          [R] 1_preprocessing.py:50-56; _advanced.py:99-105,125-128
power_to_db             [R] 1_preprocessing.py:57; _advanced.py:106,129
mfcc                    [R] 1_preprocessing.py:63-69
spectral_centroid       [R] 1_preprocessing.py:75; _advanced.py:133
spectral_bandwidth      [R] 1_preprocessing.py:77; _advanced.py:134
spectral_rolloff        [R] 1_preprocessing.py:79; _advanced.py:135
zero_crossing_rate      [R] 1_preprocessing.py:81; _advanced.py:136
rms                     [R] 1_preprocessing.py:83; _advanced.py:137
chroma_stft             [R] 1_preprocessing.py:96-101; _advanced.py:139-141
=====================  ========================================================
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.fftpack
import scipy.signal


class ParameterError(ValueError):
    """Stand-in for librosa.util.exceptions.ParameterError."""


# --------------------------------------------------------------------------
# util
# --------------------------------------------------------------------------
def tiny(x):
    x = np.asarray(x)
    if np.issubdtype(x.dtype, np.floating) or np.issubdtype(x.dtype, np.complexfloating):
        dtype = x.dtype
    else:
        dtype = np.dtype(np.float32)
    return np.finfo(dtype).tiny


def valid_audio(y):
    if not isinstance(y, np.ndarray):
        raise ParameterError("Audio data must be of type numpy.ndarray")
    if not np.issubdtype(y.dtype, np.floating):
        raise ParameterError("Audio data must be floating-point")
    if y.ndim == 0:
        raise ParameterError("Audio data must be at least one-dimensional")
    if not np.isfinite(y).all():
        raise ParameterError("Audio buffer is not finite everywhere")
    return True


def pad_center(data, *, size, axis=-1):
    n = data.shape[axis]
    lpad = int((size - n) // 2)
    lengths = [(0, 0)] * data.ndim
    lengths[axis] = (lpad, int(size - n - lpad))
    if lpad < 0:
        raise ParameterError(f"Target size ({size}) must be at least input size ({n})")
    return np.pad(data, lengths, mode="constant")


def frame(x, *, frame_length, hop_length):
    """librosa.util.frame(axis=-1): (..., n) -> (..., frame_length, n_frames)."""
    x = np.asarray(x)
    if x.shape[-1] < frame_length:
        raise ParameterError(f"Input is too short (n={x.shape[-1]}) for frame_length={frame_length}")
    if hop_length < 1:
        raise ParameterError(f"Invalid hop_length: {hop_length}")
    xw = np.lib.stride_tricks.sliding_window_view(x, frame_length, axis=-1)
    # xw: (..., n - frame_length + 1, frame_length) -> take every hop-th, move frame axis last
    xw = xw[..., ::hop_length, :]
    return np.moveaxis(xw, -1, -2)


def normalize(S, *, norm=np.inf, axis=0, threshold=None, fill=None):
    if threshold is None:
        threshold = tiny(S)
    elif threshold <= 0:
        raise ParameterError(f"threshold={threshold} must be strictly positive")
    if not np.isfinite(S).all():
        raise ParameterError("Input must be finite")
    mag = np.abs(S).astype(float)
    if norm is None:
        return S
    elif norm == np.inf:
        length = np.max(mag, axis=axis, keepdims=True)
    elif norm == -np.inf:
        length = np.min(mag, axis=axis, keepdims=True)
    elif norm == 0:
        length = np.sum(mag > 0, axis=axis, keepdims=True, dtype=mag.dtype)
    elif np.issubdtype(type(norm), np.number) and norm > 0:
        length = np.sum(mag**norm, axis=axis, keepdims=True) ** (1.0 / norm)
    else:
        raise ParameterError(f"Unsupported norm: {norm!r}")
    small_idx = length < threshold
    Snorm = np.empty_like(S)
    if fill is None:
        length[small_idx] = 1.0
        Snorm[:] = S / length
    elif fill:
        length[small_idx] = np.nan
        Snorm[:] = S / length
        Snorm[np.isnan(Snorm)] = 1.0
    else:
        length[small_idx] = np.inf
        Snorm[:] = S / length
    return Snorm


# --------------------------------------------------------------------------
# frequencies / filterbanks
# --------------------------------------------------------------------------
def fft_frequencies(*, sr=22050, n_fft=2048):
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def hz_to_mel(frequencies, *, htk=False):
    frequencies = np.asanyarray(frequencies)
    if htk:
        return 2595.0 * np.log10(1.0 + frequencies / 700.0)
    f_min = 0.0
    f_sp = 200.0 / 3
    mels = (frequencies - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if frequencies.ndim:
        log_t = frequencies >= min_log_hz
        mels[log_t] = min_log_mel + np.log(frequencies[log_t] / min_log_hz) / logstep
    elif frequencies >= min_log_hz:
        mels = min_log_mel + np.log(frequencies / min_log_hz) / logstep
    return mels


def mel_to_hz(mels, *, htk=False):
    mels = np.asanyarray(mels)
    if htk:
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_min = 0.0
    f_sp = 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_frequencies(n_mels=128, *, fmin=0.0, fmax=11025.0, htk=False):
    min_mel = hz_to_mel(fmin, htk=htk)
    max_mel = hz_to_mel(fmax, htk=htk)
    mels = np.linspace(min_mel, max_mel, n_mels)
    return mel_to_hz(mels, htk=htk)


def mel(*, sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney", dtype=np.float32):
    """librosa.filters.mel."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    fftfreqs = fft_frequencies(sr=sr, n_fft=n_fft)
    mel_f = mel_frequencies(n_mels + 2, fmin=fmin, fmax=fmax, htk=htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if isinstance(norm, str):
        if norm == "slaney":
            enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
            weights *= enorm[:, np.newaxis]
        else:
            raise ParameterError(f"Unsupported norm={norm}")
    elif norm is not None:
        weights = normalize(weights, norm=norm, axis=-1)
    return weights


# --------------------------------------------------------------------------
# stft and spectrogram
# --------------------------------------------------------------------------
def get_window(window, Nx, *, fftbins=True):
    if callable(window):
        return window(Nx)
    if isinstance(window, (str, tuple)) or np.isscalar(window):
        return scipy.signal.get_window(window, Nx, fftbins=fftbins)
    if isinstance(window, (np.ndarray, list)):
        if len(window) == Nx:
            return np.asarray(window)
        raise ParameterError(f"Window size mismatch: {len(window)} != {Nx}")
    raise ParameterError(f"Invalid window specification: {window!r}")


def num_frames(n, *, n_fft=2048, hop_length=512, center=True):
    if center:
        n = n + 2 * (n_fft // 2)
    if n < n_fft:
        raise ParameterError(f"Input is too short (n={n}) for frame_length={n_fft}")
    return 1 + (n - n_fft) // hop_length


def stft(y, *, n_fft=2048, hop_length=None, win_length=None, window="hann",
         center=True, dtype=None, pad_mode="constant"):
    """librosa.stft: (..., n) f32 -> (..., 1+n_fft/2, T) complex64."""
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    elif not (isinstance(hop_length, (int, np.integer)) and hop_length > 0):
        raise ParameterError(f"hop_length={hop_length} must be a positive integer")
    valid_audio(y)
    fft_window = get_window(window, win_length, fftbins=True)
    fft_window = pad_center(fft_window, size=n_fft)
    fft_window = fft_window.reshape((-1, 1))  # broadcast over frames
    if center:
        if pad_mode in ("wrap", "maximum", "mean", "median", "minimum"):
            raise ParameterError(f"pad_mode='{pad_mode}' is not supported by librosa.stft")
        padding = [(0, 0)] * y.ndim
        padding[-1] = (n_fft // 2, n_fft // 2)
        y = np.pad(y, padding, mode=pad_mode)
    elif n_fft > y.shape[-1]:
        raise ParameterError(f"n_fft={n_fft} is too large for uncentered analysis of input signal of length={y.shape[-1]}")
    y_frames = frame(y, frame_length=n_fft, hop_length=hop_length)
    if dtype is None:
        dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    # window is float64 -> product float64 -> float64 FFT -> cast (librosa preallocates `dtype`)
    out = scipy.fft.rfft(fft_window * y_frames, axis=-2)
    return out.astype(dtype)


def _spectrogram(*, y=None, S=None, n_fft=2048, hop_length=512, power=1, win_length=None,
                 window="hann", center=True, pad_mode="constant"):
    if S is not None:
        if n_fft is None or n_fft // 2 + 1 != S.shape[-2]:
            n_fft = 2 * (S.shape[-2] - 1)
    else:
        if n_fft is None:
            raise ParameterError(f"Unable to compute spectrogram with n_fft={n_fft}")
        if y is None:
            raise ParameterError("Input signal must be provided to compute a spectrogram")
        S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                        center=center, window=window, pad_mode=pad_mode)) ** power
    return S, n_fft


def melspectrogram(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None,
                   window="hann", center=True, pad_mode="constant", power=2.0, **kwargs):
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, power=power,
                            win_length=win_length, window=window, center=center, pad_mode=pad_mode)
    mel_basis = mel(sr=sr, n_fft=n_fft, **kwargs)
    return np.einsum("...ft,mf->...mt", S, mel_basis, optimize=True)


def power_to_db(S, *, ref=1.0, amin=1e-10, top_db=80.0):
    """librosa.power_to_db. ``ref`` may be callable (np.max)."""
    S = np.asarray(S)
    if amin <= 0:
        raise ParameterError("amin must be strictly positive")
    if np.issubdtype(S.dtype, np.complexfloating):
        magnitude = np.abs(S)
    else:
        magnitude = S
    if callable(ref):
        ref_value = ref(magnitude)
    else:
        ref_value = np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        if top_db < 0:
            raise ParameterError("top_db must be non-negative")
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def mfcc(*, y=None, sr=22050, S=None, n_mfcc=20, dct_type=2, norm="ortho", lifter=0, **kwargs):
    """librosa.feature.mfcc."""
    if S is None:
        S = power_to_db(melspectrogram(y=y, sr=sr, **kwargs))
    M = scipy.fftpack.dct(S, axis=-2, type=dct_type, norm=norm)[..., :n_mfcc, :]
    if lifter > 0:
        LI = np.sin(np.pi * np.arange(1, 1 + n_mfcc, dtype=M.dtype) / lifter)
        LI = LI.reshape((-1, 1))
        M *= 1 + (lifter / 2) * LI
        return M
    elif lifter == 0:
        return M
    raise ParameterError(f"MFCC lifter={lifter} must be a non-negative number")


# --------------------------------------------------------------------------
# spectral statistics
# --------------------------------------------------------------------------
def spectral_centroid(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, freq=None,
                      win_length=None, window="hann", center=True, pad_mode="constant"):
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                            window=window, center=center, pad_mode=pad_mode)
    if not np.isrealobj(S):
        raise ParameterError("Spectral centroid is only defined with real-valued input")
    elif np.any(S < 0):
        raise ParameterError("Spectral centroid is only defined with non-negative energies")
    if freq is None:
        freq = fft_frequencies(sr=sr, n_fft=n_fft)
    if freq.ndim == 1:
        freq = freq.reshape((-1, 1))
    return np.sum(freq * normalize(S, norm=1, axis=-2), axis=-2, keepdims=True)


def spectral_bandwidth(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None,
                       window="hann", center=True, pad_mode="constant", freq=None, centroid=None,
                       norm=True, p=2):
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                            window=window, center=center, pad_mode=pad_mode)
    if not np.isrealobj(S):
        raise ParameterError("Spectral bandwidth is only defined with real-valued input")
    elif np.any(S < 0):
        raise ParameterError("Spectral bandwidth is only defined with non-negative energies")
    if centroid is None:
        centroid = spectral_centroid(y=y, sr=sr, S=S, n_fft=n_fft, hop_length=hop_length, freq=freq)
    if freq is None:
        freq = fft_frequencies(sr=sr, n_fft=n_fft)
    if freq.ndim == 1:
        deviation = np.abs(np.subtract.outer(centroid[..., 0, :], freq).swapaxes(-2, -1))
    else:
        deviation = np.abs(freq - centroid)
    if norm:
        S = normalize(S, norm=1, axis=-2)
    return np.sum(S * deviation**p, axis=-2, keepdims=True) ** (1.0 / p)


def spectral_rolloff(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None,
                     window="hann", center=True, pad_mode="constant", freq=None, roll_percent=0.85):
    if not 0.0 < roll_percent < 1.0:
        raise ParameterError("roll_percent must lie in the range (0, 1)")
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                            window=window, center=center, pad_mode=pad_mode)
    if not np.isrealobj(S):
        raise ParameterError("Spectral rolloff is only defined with real-valued input")
    elif np.any(S < 0):
        raise ParameterError("Spectral rolloff is only defined with non-negative energies")
    if freq is None:
        freq = fft_frequencies(sr=sr, n_fft=n_fft)
    if freq.ndim == 1:
        freq = freq.reshape((-1, 1))
    total_energy = np.cumsum(S, axis=-2)
    threshold = roll_percent * total_energy[..., -1, :]
    threshold = np.expand_dims(threshold, axis=-2)
    ind = np.where(total_energy < threshold, np.nan, 1)
    return np.nanmin(ind * freq, axis=-2, keepdims=True)


def zero_crossings(y, *, threshold=1e-10, ref_magnitude=None, pad=True, zero_pos=True, axis=-1):
    if threshold is None:
        threshold = 0.0
    if callable(ref_magnitude):
        threshold = threshold * ref_magnitude(np.abs(y))
    elif ref_magnitude is not None:
        threshold = threshold * ref_magnitude
    yc = np.where(np.abs(y) <= threshold, 0, y) if threshold >= 0 else y
    if zero_pos:
        s = np.signbit(yc)
    else:
        s = np.sign(yc)
    s = np.moveaxis(s, axis, -1)
    z = np.empty(s.shape, dtype=bool)
    z[..., 0] = pad
    z[..., 1:] = s[..., 1:] != s[..., :-1]
    return np.moveaxis(z, -1, axis)


def zero_crossing_rate(y, *, frame_length=2048, hop_length=512, center=True, **kwargs):
    valid_audio(y)
    if center:
        padding = [(0, 0)] * y.ndim
        padding[-1] = (frame_length // 2, frame_length // 2)
        y = np.pad(y, padding, mode="edge")
    y_framed = frame(y, frame_length=frame_length, hop_length=hop_length)
    kwargs["axis"] = -2
    kwargs.setdefault("pad", False)
    crossings = zero_crossings(y_framed, **kwargs)
    return np.mean(crossings, axis=-2, keepdims=True)


def rms(*, y=None, S=None, frame_length=2048, hop_length=512, center=True, pad_mode="constant",
        dtype=np.float32):
    if y is not None:
        if center:
            padding = [(0, 0)] * y.ndim
            padding[-1] = (frame_length // 2, frame_length // 2)
            y = np.pad(y, padding, mode=pad_mode)
        x = frame(y, frame_length=frame_length, hop_length=hop_length)
        power = np.mean(np.square(x.astype(dtype, copy=False), dtype=dtype), axis=-2, keepdims=True)
    elif S is not None:
        if S.shape[-2] != frame_length // 2 + 1:
            raise ParameterError("Since S.shape[-2] is {}, frame_length is expected to be {} or {}".format(
                S.shape[-2], S.shape[-2] * 2 - 2, S.shape[-2] * 2 - 1))
        x = np.square(np.abs(S).astype(dtype, copy=False), dtype=dtype)
        x[..., 0, :] *= 0.5
        if frame_length % 2 == 0:
            x[..., -1, :] *= 0.5
        power = 2 * np.sum(x, axis=-2, keepdims=True) / frame_length**2
    else:
        raise ParameterError("Either `y` or `S` must be input.")
    return np.sqrt(power)


# --------------------------------------------------------------------------
# chroma_stft ("next" row)
# --------------------------------------------------------------------------
def _parabolic_interpolation(x, *, axis=-2):
    x = np.moveaxis(x, axis, 0)
    shifts = np.zeros_like(x)
    a = x[2:] + x[:-2] - 2 * x[1:-1]
    b = (x[2:] - x[:-2]) / 2
    with np.errstate(divide="ignore", invalid="ignore"):
        s = -b / a
    s[np.abs(b) >= np.abs(a)] = 0
    shifts[1:-1] = s
    return np.moveaxis(shifts, 0, axis)


def _localmax(x, *, axis=0):
    x = np.moveaxis(x, axis, 0)
    out = np.zeros(x.shape, dtype=bool)
    out[1:-1] = (x[1:-1] > x[:-2]) & (x[1:-1] >= x[2:])
    out[-1] = x[-1] > x[-2]
    return np.moveaxis(out, 0, axis)


def piptrack(*, S, sr=22050, n_fft=2048, fmin=150.0, fmax=4000.0, threshold=0.1, ref=None):
    S = np.abs(S)
    fmin = np.maximum(fmin, 0)
    fmax = np.minimum(fmax, float(sr) / 2)
    fft_freqs = fft_frequencies(sr=sr, n_fft=n_fft)
    avg = np.gradient(S, axis=-2)
    shift = _parabolic_interpolation(S, axis=-2)
    dskew = 0.5 * avg * shift
    pitches = np.zeros_like(S)
    mags = np.zeros_like(S)
    freq_mask = (fmin <= fft_freqs) & (fft_freqs < fmax)
    freq_mask = freq_mask.reshape((-1, 1))
    if ref is None:
        ref = np.max
    if callable(ref):
        ref_value = threshold * ref(S, axis=-2)
        ref_value = np.expand_dims(ref_value, -2)
    else:
        ref_value = np.abs(ref)
    idx = np.nonzero(freq_mask & _localmax(S * (S > ref_value), axis=-2))
    pitches[idx] = (idx[-2] + shift[idx]) * float(sr) / n_fft
    mags[idx] = S[idx] + dskew[idx]
    return pitches, mags


def pitch_tuning(frequencies, *, resolution=0.01, bins_per_octave=12):
    frequencies = np.atleast_1d(frequencies)
    frequencies = frequencies[frequencies > 0]
    if not np.any(frequencies):
        return 0.0
    # hz_to_octs(f, tuning=0, bins_per_octave) = log2(f / (440/16))
    residual = np.mod(bins_per_octave * np.log2(frequencies / (440.0 / 16)), 1.0)
    residual[residual >= 0.5] -= 1.0
    bins = np.linspace(-0.5, 0.5, int(np.ceil(1.0 / resolution)) + 1)
    counts, tuning = np.histogram(residual, bins)
    return float(tuning[np.argmax(counts)])


def estimate_tuning(*, S, sr=22050, n_fft=2048, resolution=0.01, bins_per_octave=12, **kwargs):
    pitch, mag = piptrack(S=S, sr=sr, n_fft=n_fft, **kwargs)
    pitch_mask = pitch > 0
    if pitch_mask.any():
        threshold = np.median(mag[pitch_mask])
    else:
        threshold = 0.0
    return pitch_tuning(pitch[(mag >= threshold) & pitch_mask], resolution=resolution,
                        bins_per_octave=bins_per_octave)


def chroma_filter(*, sr, n_fft, n_chroma=12, tuning=0.0, ctroct=5.0, octwidth=2, norm=2,
                  base_c=True, dtype=np.float32):
    wts = np.zeros((n_chroma, n_fft))
    frequencies = np.linspace(0, sr, n_fft, endpoint=False)[1:]
    A440 = 440.0 * 2.0 ** (tuning / n_chroma)
    frqbins = n_chroma * np.log2(frequencies / (A440 / 16))
    frqbins = np.concatenate(([frqbins[0] - 1.5 * n_chroma], frqbins))
    binwidthbins = np.concatenate((np.maximum(frqbins[1:] - frqbins[:-1], 1.0), [1]))
    D = np.subtract.outer(frqbins, np.arange(0, n_chroma, dtype="d")).T
    n_chroma2 = np.round(float(n_chroma) / 2)
    D = np.remainder(D + n_chroma2 + 10 * n_chroma, n_chroma) - n_chroma2
    wts = np.exp(-0.5 * (2 * D / np.tile(binwidthbins, (n_chroma, 1))) ** 2)
    wts = normalize(wts, norm=norm, axis=0)
    if octwidth is not None:
        wts *= np.tile(np.exp(-0.5 * (((frqbins / n_chroma - ctroct) / octwidth) ** 2)), (n_chroma, 1))
    if base_c:
        wts = np.roll(wts, -3 * (n_chroma // 12), axis=0)
    return np.ascontiguousarray(wts[:, : int(1 + n_fft / 2)], dtype=dtype)


def chroma_stft(*, y=None, sr=22050, S=None, norm=np.inf, n_fft=2048, hop_length=512,
                win_length=None, window="hann", center=True, pad_mode="constant", tuning=None,
                n_chroma=12, **kwargs):
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, power=2,
                            win_length=win_length, window=window, center=center, pad_mode=pad_mode)
    if tuning is None:
        tuning = estimate_tuning(S=S, sr=sr, n_fft=n_fft, bins_per_octave=n_chroma)
    chromafb = chroma_filter(sr=sr, n_fft=n_fft, tuning=tuning, n_chroma=n_chroma, **kwargs)
    raw_chroma = np.einsum("cf,...ft->...ct", chromafb, S, optimize=True)
    return normalize(raw_chroma, norm=norm, axis=-2)


# --------------------------------------------------------------------------
# script-level helpers (the reference's own functions, restated)
# --------------------------------------------------------------------------
BASIC_CONFIG = dict(sample_rate=22050, duration=30, n_mels=128, n_fft=2048, hop_length=512, n_mfcc=40)
ADV_CONFIG = dict(sample_rate=22050, duration=30, n_mels=128, n_fft=2048, hop_length=512,
                  fixed_time_steps=1024)


def basic_extract_mel_spectrogram(audio, sr, cfg=BASIC_CONFIG):
    """[R] src/1_preprocessing.py:48-58."""
    m = melspectrogram(y=audio, sr=sr, n_mels=cfg["n_mels"], n_fft=cfg["n_fft"],
                       hop_length=cfg["hop_length"])
    return power_to_db(m, ref=np.max)


def basic_extract_mfcc(audio, sr, cfg=BASIC_CONFIG):
    """[R] src/1_preprocessing.py:61-70."""
    return mfcc(y=audio, sr=sr, n_mfcc=cfg["n_mfcc"], n_fft=cfg["n_fft"],
                hop_length=cfg["hop_length"])


def extract_spectral_features(audio, sr, cfg=BASIC_CONFIG):
    """[R] src/1_preprocessing.py:73-91 (n_fft is NOT forwarded there: librosa default 2048)."""
    return {
        "spectral_centroid": spectral_centroid(y=audio, sr=sr, hop_length=cfg["hop_length"]),
        "spectral_bandwidth": spectral_bandwidth(y=audio, sr=sr, hop_length=cfg["hop_length"]),
        "spectral_rolloff": spectral_rolloff(y=audio, sr=sr, hop_length=cfg["hop_length"]),
        "zcr": zero_crossing_rate(audio, hop_length=cfg["hop_length"]),
        "rms": rms(y=audio, hop_length=cfg["hop_length"]),
    }


def extract_chroma_features(audio, sr, cfg=BASIC_CONFIG):
    """[R] src/1_preprocessing.py:94-102."""
    return chroma_stft(y=audio, sr=sr, n_fft=cfg["n_fft"], hop_length=cfg["hop_length"])


def extract_all_features(audio, sr, cfg=BASIC_CONFIG, with_chroma=True):
    """[R] src/1_preprocessing.py:105-129 -> (370,) float64 (346 when with_chroma=False)."""
    mel_spec = basic_extract_mel_spectrogram(audio, sr, cfg)
    mf = basic_extract_mfcc(audio, sr, cfg)
    spectral = extract_spectral_features(audio, sr, cfg)
    features = []
    features.extend(np.mean(mel_spec, axis=1))
    features.extend(np.std(mel_spec, axis=1))
    features.extend(np.mean(mf, axis=1))
    features.extend(np.std(mf, axis=1))
    for _name, feat in spectral.items():
        features.append(np.mean(feat))
        features.append(np.std(feat))
    if with_chroma:
        chroma = extract_chroma_features(audio, sr, cfg)
        features.extend(np.mean(chroma, axis=1))
        features.extend(np.std(chroma, axis=1))
    return np.array(features)


def adv_extract_mel_spectrogram(audio, sr, cfg=ADV_CONFIG):
    """[R] src/1_preprocessing_advanced.py:97-114 -> (n_mels, fixed_time_steps) f32."""
    m = melspectrogram(y=audio, sr=sr, n_mels=cfg["n_mels"], n_fft=cfg["n_fft"],
                       hop_length=cfg["hop_length"])
    mel_db = power_to_db(m, ref=np.max)
    fts = cfg["fixed_time_steps"]
    if mel_db.shape[1] > fts:
        mel_db = mel_db[:, :fts]
    else:
        pad_width = fts - mel_db.shape[1]
        mel_db = np.pad(mel_db, ((0, 0), (0, pad_width)), mode="constant",
                        constant_values=mel_db.min())
    return mel_db


def extract_flattened_features(audio, sr, cfg=ADV_CONFIG, with_chroma=True):
    """[R] src/1_preprocessing_advanced.py:120-156 -> (290,) float64 (266 without chroma)."""
    m = melspectrogram(y=audio, sr=sr, n_mels=128, n_fft=cfg["n_fft"], hop_length=cfg["hop_length"])
    mel_db = power_to_db(m, ref=np.max)
    feats = [
        spectral_centroid(y=audio, sr=sr, hop_length=cfg["hop_length"]),
        spectral_bandwidth(y=audio, sr=sr, hop_length=cfg["hop_length"]),
        spectral_rolloff(y=audio, sr=sr, hop_length=cfg["hop_length"]),
        zero_crossing_rate(audio, hop_length=cfg["hop_length"]),
        rms(y=audio, hop_length=cfg["hop_length"]),
    ]
    features = []
    features.extend(np.mean(mel_db, axis=1))
    features.extend(np.std(mel_db, axis=1))
    for feat in feats:
        features.append(np.mean(feat))
        features.append(np.std(feat))
    if with_chroma:
        chroma = chroma_stft(y=audio, sr=sr, n_fft=cfg["n_fft"], hop_length=cfg["hop_length"])
        features.extend(np.mean(chroma, axis=1))
        features.extend(np.std(chroma, axis=1))
    return np.array(features)


# --------------------------------------------------------------------------
# librosa.load's arithmetic (the scripts' load_audio_file, [R] src/1_preprocessing.py:137-153,
# src/1_preprocessing_advanced.py:79-94), restated for PCM16 frames that the caller has already read [L]:
#   soundfile read(dtype=float32): int16 / 32768            -> pcm16_to_float
#   librosa.to_mono: np.mean over the channel axis           -> to_mono
#   librosa.resample(res_type="polyphase"): scipy.signal.resample_poly(y, sr/gcd, orig/gcd) followed by
#     util.fix_length(size=ceil(n * ratio))                  -> resample
# librosa.load's DEFAULT res_type is "soxr_hq" (the soxr library, not installed here and a different
# low-pass): the device front end implements the "polyphase" type, which IS scipy's resample_poly, so
# this part of the oracle is pinned by a third-party implementation.
# --------------------------------------------------------------------------
def pcm16_to_float(frames):
    return (np.asarray(frames, dtype=np.int16).astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def to_mono(y):
    """y: (channels, n) as soundfile hands it to librosa (after the transpose)."""
    if y.ndim > 1:
        y = np.mean(y, axis=tuple(range(y.ndim - 1)))
    return y


def fix_length(data, *, size, axis=-1):
    n = data.shape[axis]
    if n > size:
        slices = [slice(None)] * data.ndim
        slices[axis] = slice(0, size)
        return data[tuple(slices)]
    if n < size:
        lengths = [(0, 0)] * data.ndim
        lengths[axis] = (0, size - n)
        return np.pad(data, lengths, mode="constant")
    return data


def resample(y, *, orig_sr, target_sr, res_type="polyphase", fix=True):
    if orig_sr == target_sr:
        return y
    if res_type != "polyphase":
        raise ParameterError("the oracle restates res_type='polyphase' only (soxr is not installed)")
    ratio = float(target_sr) / orig_sr
    n_samples = int(np.ceil(y.shape[-1] * ratio))
    g = np.gcd(int(orig_sr), int(target_sr))
    y_hat = scipy.signal.resample_poly(y, int(target_sr) // g, int(orig_sr) // g, axis=-1)
    if fix:
        y_hat = fix_length(y_hat, size=n_samples)
    return np.asarray(y_hat, dtype=y.dtype)


def load_pcm16(frames, sr_native, *, sr=22050, mono=True, duration=None, res_type="polyphase"):
    """librosa.load for PCM16 frames (n, channels) already read from the container -> (y float32, sr)."""
    frames = np.asarray(frames)
    if frames.ndim == 1:
        frames = frames[:, None]
    if duration is not None:
        frames = frames[: int(duration * sr_native)]
    y = pcm16_to_float(frames).T                     # (channels, n)
    if mono:
        y = to_mono(y)
    elif y.shape[0] == 1:
        y = y[0]
    if sr is not None and sr != sr_native:
        y = resample(y, orig_sr=sr_native, target_sr=sr, res_type=res_type)
    else:
        sr = sr_native
    return np.asarray(y, dtype=np.float32), sr


def load_audio_file_pcm16(frames, sr_native, cfg=BASIC_CONFIG):
    """[R] src/1_preprocessing.py:137-153 after the container is read: load + right zero pad."""
    audio, sr = load_pcm16(frames, sr_native, sr=cfg["sample_rate"], duration=cfg["duration"])
    expected = cfg["sample_rate"] * cfg["duration"]
    if len(audio) < expected:
        audio = np.pad(audio, (0, expected - len(audio)), mode="constant")
    return audio, sr
