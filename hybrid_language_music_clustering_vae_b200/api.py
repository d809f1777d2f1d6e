"""librosa-compatible call surface on top of the CUDA path.

Same names, keyword-only signatures, shapes ``(..., n_feat, T)`` and dtypes as
the librosa functions the reference calls; leading dimensions are a batch.
numpy in -> numpy out (through the pinned-host pipeline); CUDA tensor in ->
CUDA tensor out.  Deviation from librosa, on purpose: for batched input
``power_to_db(ref=np.max)`` and ``top_db`` act per clip (per leading index),
because the scripts call librosa once per clip.
"""
from __future__ import annotations

import numpy as np

from .core import ParameterError, UnsupportedError, get_extractor, _ref_to_mode, _check
from . import _lib


def _is_cuda(x):
    try:
        import torch

        return isinstance(x, torch.Tensor) and x.is_cuda
    except ImportError:  # pragma: no cover
        return False


def _prep(y):
    """-> (2-D batch, leading shape, on_device)."""
    if y is None:
        raise ParameterError("Input signal must be provided")
    if _is_cuda(y):
        lead = tuple(y.shape[:-1])
        return y.reshape(-1, y.shape[-1]), lead, True
    try:
        import torch

        if isinstance(y, torch.Tensor):
            y = y.numpy()
    except ImportError:  # pragma: no cover
        pass
    if not isinstance(y, np.ndarray):
        raise ParameterError("Audio data must be of type numpy.ndarray")
    if not np.issubdtype(y.dtype, np.floating):
        raise ParameterError("Audio data must be floating-point")
    if y.ndim == 0:
        raise ParameterError("Audio data must be at least one-dimensional")
    lead = tuple(y.shape[:-1])
    return y.reshape(-1, y.shape[-1]), lead, False


def _valid_or_raise(status):
    """librosa.util.valid_audio: non-finite audio raises."""
    bad = np.asarray(status if isinstance(status, np.ndarray) else status.cpu().numpy())
    if bad.any():
        raise ParameterError("Audio buffer is not finite everywhere")


def _device_of(y, on_dev):
    return y.device.index if on_dev else 0


def _stft_kw(n_fft, hop_length, win_length, window, center, pad_mode):
    return dict(n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window,
                center=center, pad_mode=pad_mode)


def stft(y, *, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True, dtype=None,
         pad_mode="constant", out=None):
    """librosa.stft -> (..., 1 + n_fft/2, T) complex64."""
    yb, lead, on_dev = _prep(y)
    import torch

    ex = get_extractor(**_stft_kw(n_fft, hop_length, win_length, window, center, pad_mode), n_mfcc=0,
                       device=_device_of(yb, on_dev))
    if on_dev:
        spec = ex.stft_device(yb)
        return spec.reshape(lead + tuple(spec.shape[1:]))
    if not np.isfinite(yb).all():
        raise ParameterError("Audio buffer is not finite everywhere")
    dev = torch.device("cuda", ex.device)
    spec = ex.stft_device(torch.from_numpy(np.ascontiguousarray(yb, dtype=np.float32)).to(dev)).cpu().numpy()
    return spec.reshape(lead + spec.shape[1:])


def power_to_db(S, *, ref=1.0, amin=1e-10, top_db=80.0):
    """librosa.power_to_db; ref=np.max / top_db per leading index for >2-D input."""
    import ctypes as C
    import torch

    if amin <= 0:
        raise ParameterError("amin must be strictly positive")
    if top_db is not None and top_db < 0:
        raise ParameterError("top_db must be non-negative")
    mode, val = _ref_to_mode(ref)
    on_dev = _is_cuda(S)
    if on_dev:
        St = S if S.dtype == torch.float32 else S.abs().float()
    else:
        St = torch.from_numpy(np.ascontiguousarray(np.abs(np.asarray(S)), dtype=np.float32)).cuda()
    shape = tuple(St.shape)
    if St.dim() <= 2:
        B, rows, T = 1, 1, St.numel()
    else:
        B, rows, T = int(np.prod(shape[:-2])), shape[-2], shape[-1]
    St = St.contiguous()
    outp = torch.empty_like(St)
    cm = torch.empty((max(B, 1),), dtype=torch.float32, device=St.device)
    stream = torch.cuda.current_stream(St.device).cuda_stream
    _check(_lib.lib.hlmc_power_to_db_device(
        C.c_void_p(St.data_ptr()), C.c_void_p(outp.data_ptr()), B, rows, T, mode, val, float(amin),
        float(top_db) if top_db is not None else -1.0, C.c_void_p(cm.data_ptr()), St.device.index,
        C.c_void_p(stream)))
    return outp if on_dev else outp.cpu().numpy()


class _Feature:
    """Namespace mirroring ``librosa.feature``."""

    @staticmethod
    def _run(y, kw, *, want):
        yb, lead, on_dev = _prep(y)
        ex = get_extractor(**kw, device=_device_of(yb, on_dev))
        if on_dev:
            # device input: asynchronous, no host synchronisation - non-finite audio is NOT raised here (that
            # would need a device-to-host read); use FeatureExtractor.extract_device and inspect ``status``
            res = ex.extract_device(yb, mfcc=(want == "mfcc"), stats=(want == "stats"))
        else:
            if not np.isfinite(yb).all():          # librosa.util.valid_audio: the whole buffer, not only framed samples
                raise ParameterError("Audio buffer is not finite everywhere")
            res = ex.extract_host(yb, logmel=(want in ("logmel",)), mfcc=(want == "mfcc"),
                                  stats=(want == "stats"))
            _valid_or_raise(res["status"])
        return res, lead

    @staticmethod
    def melspectrogram(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None,
                       window="hann", center=True, pad_mode="constant", power=2.0, **kwargs):
        """librosa.feature.melspectrogram -> (..., n_mels, T) float32 mel power."""
        if S is not None:
            raise UnsupportedError("melspectrogram(S=...) is not implemented on device; pass y")
        import torch

        yb, lead, on_dev = _prep(y)
        ex = get_extractor(sr=sr, **_stft_kw(n_fft, hop_length, win_length, window, center, pad_mode),
                           power=power, n_mfcc=0, device=_device_of(yb, on_dev), **kwargs)
        if on_dev:
            mel = ex.melspectrogram_device(yb)
            return mel.reshape(lead + tuple(mel.shape[1:]))
        if not np.isfinite(yb).all():
            raise ParameterError("Audio buffer is not finite everywhere")
        dev = torch.device("cuda", ex.device)
        mel = ex.melspectrogram_device(torch.from_numpy(np.ascontiguousarray(yb, dtype=np.float32)).to(dev))
        mel = mel.cpu().numpy()
        return mel.reshape(lead + mel.shape[1:])

    @staticmethod
    def mfcc(*, y=None, sr=22050, S=None, n_mfcc=20, dct_type=2, norm="ortho", lifter=0, **kwargs):
        """librosa.feature.mfcc -> (..., n_mfcc, T) float32."""
        if S is not None:
            raise UnsupportedError("mfcc(S=...) is not implemented on device; pass y")
        if dct_type != 2 or norm != "ortho":
            raise UnsupportedError("only dct_type=2, norm='ortho' is implemented")
        if lifter < 0:
            raise ParameterError(f"MFCC lifter={lifter} must be a non-negative number")
        mel_kw = dict(kwargs)
        mel_norm = mel_kw.pop("mel_norm", "slaney")   # librosa.filters.mel's `norm` cannot be passed via mfcc
        res, lead = _Feature._run(y, dict(sr=sr, n_mfcc=n_mfcc, lifter=lifter, norm=mel_norm, **mel_kw),
                                  want="mfcc")
        m = res["mfcc"]
        return m.reshape(lead + tuple(m.shape[1:]))

    @staticmethod
    def chroma_stft(*, y=None, sr=22050, S=None, norm=np.inf, n_fft=2048, hop_length=512, win_length=None,
                    window="hann", center=True, pad_mode="constant", tuning=None, n_chroma=12, **kwargs):
        """librosa.feature.chroma_stft -> (..., 12, T) float32, tuning estimated per clip."""
        import torch

        if S is not None or tuning is not None or n_chroma != 12 or norm != np.inf or kwargs:
            raise UnsupportedError("chroma_stft: only y input with the default tuning=None, n_chroma=12, "
                                   "norm=inf and filterbank parameters is implemented on device")
        yb, lead, on_dev = _prep(y)
        ex = get_extractor(sr=sr, **_stft_kw(n_fft, hop_length, win_length, window, center, pad_mode), n_mfcc=0,
                           device=_device_of(yb, on_dev))
        if on_dev:
            c = ex.extract_device(yb, mfcc=False, stats=False, chroma=True)["chroma"]
            return c.reshape(lead + tuple(c.shape[1:]))
        if not np.isfinite(yb).all():
            raise ParameterError("Audio buffer is not finite everywhere")
        dev = torch.device("cuda", ex.device)
        c = ex.extract_device(torch.from_numpy(np.ascontiguousarray(yb, dtype=np.float32)).to(dev), mfcc=False,
                              stats=False, chroma=True)["chroma"].cpu().numpy()
        return c.reshape(lead + c.shape[1:])

    @staticmethod
    def _stat(idx, y, kw, dtype64=True):
        res, lead = _Feature._run(y, dict(n_mfcc=0, **kw), want="stats")
        s = res["stats"][:, idx:idx + 1, :]
        s = s.reshape(lead + tuple(s.shape[1:]))
        if isinstance(s, np.ndarray):
            return s.astype(np.float64) if dtype64 else np.ascontiguousarray(s)
        return s.double() if dtype64 else s.contiguous()

    @staticmethod
    def spectral_centroid(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, freq=None,
                          win_length=None, window="hann", center=True, pad_mode="constant"):
        if S is not None or freq is not None:
            raise UnsupportedError("spectral_centroid(S=/freq=) is not implemented on device")
        return _Feature._stat(0, y, dict(sr=sr, **_stft_kw(n_fft, hop_length, win_length, window, center, pad_mode)))

    @staticmethod
    def spectral_bandwidth(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None,
                           window="hann", center=True, pad_mode="constant", freq=None, centroid=None,
                           norm=True, p=2):
        if S is not None or freq is not None or centroid is not None or not norm or p != 2:
            raise UnsupportedError("spectral_bandwidth: only y input, norm=True, p=2 are implemented")
        return _Feature._stat(1, y, dict(sr=sr, **_stft_kw(n_fft, hop_length, win_length, window, center, pad_mode)))

    @staticmethod
    def spectral_rolloff(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None,
                         window="hann", center=True, pad_mode="constant", freq=None, roll_percent=0.85):
        if not 0.0 < roll_percent < 1.0:
            raise ParameterError("roll_percent must lie in the range (0, 1)")
        if S is not None or freq is not None:
            raise UnsupportedError("spectral_rolloff(S=/freq=) is not implemented on device")
        return _Feature._stat(2, y, dict(sr=sr, roll_percent=roll_percent,
                                         **_stft_kw(n_fft, hop_length, win_length, window, center, pad_mode)))

    @staticmethod
    def zero_crossing_rate(y, *, frame_length=2048, hop_length=512, center=True, **kwargs):
        thr = kwargs.pop("threshold", 1e-10)
        if kwargs.pop("ref_magnitude", None) is not None or not kwargs.pop("zero_pos", True) or kwargs.pop("pad", False):
            raise UnsupportedError("zero_crossing_rate: only threshold= is implemented")
        thr = 0.0 if thr is None else thr
        return _Feature._stat(3, y, dict(n_fft=frame_length, hop_length=hop_length, center=center,
                                         zcr_threshold=thr))

    @staticmethod
    def rms(*, y=None, S=None, frame_length=2048, hop_length=512, center=True, pad_mode="constant",
            dtype=np.float32):
        if S is not None:
            raise UnsupportedError("rms(S=...) is not implemented on device; pass y")
        return _Feature._stat(4, y, dict(n_fft=frame_length, hop_length=hop_length, center=center,
                                         pad_mode=pad_mode), dtype64=False)


feature = _Feature()
