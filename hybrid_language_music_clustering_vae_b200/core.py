"""Host-side driver: one ``FeatureExtractor`` per parameter set and device.

It owns a C-ABI plan (``hlmc_plan``) and exposes the batched calls the two
preprocessing scripts use instead of their per-file librosa loops
([R] src/1_preprocessing.py:223-258, src/1_preprocessing_advanced.py:286-314).
torch is used only to own device memory and streams; numpy inputs go through
the library's own pinned-host pipeline (``hlmc_extract_host``).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional

import numpy as np

from . import _lib
from ._lib import HlmcHostIo, HlmcParams, lib

STAT_NAMES = ("spectral_centroid", "spectral_bandwidth", "spectral_rolloff", "zcr", "rms")


class ParameterError(ValueError):
    """Raised where librosa raises ``librosa.util.exceptions.ParameterError``."""


class UnsupportedError(NotImplementedError):
    """Valid for librosa but outside what the CUDA path implements."""


def _check(rc: int):
    if rc == _lib.HLMC_OK:
        return
    msg = _lib.last_error()
    if rc == _lib.HLMC_ERR_PARAM:
        raise ParameterError(msg)
    if rc == _lib.HLMC_ERR_UNSUPPORTED:
        raise UnsupportedError(msg)
    raise RuntimeError(f"hlmc_b200: {msg} (status {rc})")


def _resolve_window(window, win_length: int):
    """librosa.filters.get_window: name / tuple / callable / array -> float64[win_length] or None (Hann)."""
    if isinstance(window, str) and window in ("hann", "hanning"):
        return None  # built into the C library (periodic Hann)
    if callable(window):
        w = np.asarray(window(win_length), dtype=np.float64)
    elif isinstance(window, (str, tuple)) or np.isscalar(window):
        import scipy.signal  # the same helper librosa delegates to

        w = scipy.signal.get_window(window, win_length, fftbins=True).astype(np.float64)
    else:
        w = np.asarray(window, dtype=np.float64)
    if w.shape != (win_length,):
        raise ParameterError(f"Window size mismatch: {w.shape} != ({win_length},)")
    return np.ascontiguousarray(w)


def _ref_to_mode(ref):
    if callable(ref):
        if ref is np.max or ref is np.amax or ref is max:
            return _lib.REF_MAX, 1.0
        raise UnsupportedError("power_to_db: the only callable `ref` supported on device is np.max")
    return _lib.REF_VALUE, float(abs(ref))


def _row_pitch(w):
    """Row pitch in elements; a single-row tensor may carry an arbitrary stride(0)."""
    return w.shape[1] if w.shape[0] <= 1 else w.stride(0)


class DeviceGraph:
    """A captured ``extract_device`` call (``FeatureExtractor.capture_device``)."""

    def __init__(self, extractor, handle, waves, out):
        self._ex, self._h, self.waves, self.out = extractor, handle, waves, out

    def replay(self, stream=None):
        import torch

        s = stream if stream is not None else torch.cuda.current_stream(self.waves.device)
        _check(lib.hlmc_graph_launch(self._h, C.c_void_p(s.cuda_stream)))
        return self.out

    def close(self):
        if getattr(self, "_h", None):
            lib.hlmc_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FeatureExtractor:
    """Fused log-mel + MFCC + spectral statistics for batches of equal-length clips."""

    def __init__(self, *, sr=22050, n_fft=2048, hop_length=512, win_length=None, window="hann",
                 center=True, pad_mode="constant", n_mels=128, fmin=0.0, fmax=None, htk=False,
                 norm="slaney", power=2.0, n_mfcc=20, lifter=0, ref=1.0, amin=1e-10, top_db=80.0,
                 roll_percent=0.85, zcr_threshold=1e-10, mel_basis=None, device=0):
        if pad_mode not in _lib.PAD_MODES:
            if pad_mode in ("wrap", "maximum", "mean", "median", "minimum"):
                raise ParameterError(f"pad_mode='{pad_mode}' is not supported by librosa.stft")
            raise UnsupportedError(f"pad_mode='{pad_mode}' is not implemented on device")
        if norm not in ("slaney", None):
            raise UnsupportedError("mel norm must be 'slaney' or None")
        if top_db is not None and top_db < 0:
            raise ParameterError("top_db must be non-negative")
        if hop_length is None:
            hop_length = int((win_length or n_fft) // 4)
        p = HlmcParams()
        lib.hlmc_params_default(C.byref(p))
        p.sr, p.n_fft, p.hop_length = int(sr), int(n_fft), int(hop_length)
        p.win_length = int(win_length) if win_length else int(n_fft)
        p.center, p.pad_mode = int(bool(center)), _lib.PAD_MODES[pad_mode]
        p.n_mels, p.fmin = int(n_mels), float(fmin)
        p.fmax = float(fmax) if fmax is not None else -1.0
        p.htk, p.mel_norm, p.power = int(bool(htk)), (1 if norm == "slaney" else 0), float(power)
        p.n_mfcc, p.lifter = int(n_mfcc), float(lifter)
        p.ref_mode, p.ref_value = _ref_to_mode(ref)
        p.amin = float(amin)
        p.top_db = float(top_db) if top_db is not None else -1.0
        p.roll_percent, p.zcr_threshold = float(roll_percent), float(zcr_threshold)
        self.params = p
        self.device = int(device)
        self._win = _resolve_window(window, p.win_length) if p.win_length > 0 else None
        mb = None
        if mel_basis is not None:
            mb = np.ascontiguousarray(mel_basis, dtype=np.float32)
            if mb.shape != (p.n_mels, 1 + p.n_fft // 2):
                raise ParameterError("mel_basis shape mismatch")
        self._mb = mb
        handle = C.c_void_p()
        _check(lib.hlmc_plan_create(
            C.byref(p),
            self._win.ctypes.data_as(C.c_void_p) if self._win is not None else None,
            mb.ctypes.data_as(C.c_void_p) if mb is not None else None,
            self.device, C.byref(handle)))
        self._plan = handle
        self._lock = threading.Lock()

    # -- lifetime -----------------------------------------------------------
    def close(self):
        if getattr(self, "_plan", None):
            lib.hlmc_plan_destroy(self._plan)
            self._plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- introspection ------------------------------------------------------
    @property
    def n_mels(self):
        return self.params.n_mels

    @property
    def n_mfcc(self):
        return self.params.n_mfcc

    @property
    def n_bins(self):
        return 1 + self.params.n_fft // 2

    def num_frames(self, n: int) -> int:
        t = lib.hlmc_num_frames(C.byref(self.params), int(n))
        if t < 0:
            _check(int(t))
        return int(t)

    def uses_fast_path(self) -> bool:
        return bool(lib.hlmc_plan_uses_fast_path(self._plan))

    def force_generic(self, flag: bool = True):
        _check(lib.hlmc_plan_set_path(self._plan, int(bool(flag))))

    def set_path(self, path: int):
        """0 = automatic, 1 = shared-memory FFT kernel (HLMC_PATH_GENERIC), 2 = register-FFT kernel with its tables
        in shared memory instead of Tensor Memory (HLMC_PATH_FAST_SMEM_TABLES); tests cross-check the three."""
        _check(lib.hlmc_plan_set_path(self._plan, int(path)))

    def set_timing(self, enable: bool):
        _check(lib.hlmc_plan_set_timing(self._plan, int(bool(enable))))

    def read_timing(self):
        """-> (frames-kernel ms, dB+DCT-kernel ms, calls) since the last read."""
        f, d, n = C.c_double(0), C.c_double(0), C.c_int64(0)
        _check(lib.hlmc_plan_read_timing(self._plan, C.byref(f), C.byref(d), C.byref(n)))
        return float(f.value), float(d.value), int(n.value)

    def mel_basis(self) -> np.ndarray:
        out = np.empty((self.n_mels, self.n_bins), dtype=np.float32)
        _check(lib.hlmc_plan_mel_basis(self._plan, out.ctypes.data_as(C.c_void_p)))
        return out

    def dct_basis(self) -> np.ndarray:
        out = np.empty((self.n_mfcc, self.n_mels), dtype=np.float32)
        _check(lib.hlmc_plan_dct_basis(self._plan, out.ctypes.data_as(C.c_void_p)))
        return out

    def pooled_width(self, with_mfcc=True, with_chroma=False) -> int:
        return 2 * self.n_mels + (2 * self.n_mfcc if with_mfcc else 0) + 10 + (24 if with_chroma else 0)

    # -- device-resident path -----------------------------------------------
    def _as_cuda_batch(self, waves):
        import torch

        if not (isinstance(waves, torch.Tensor) and waves.is_cuda):
            raise TypeError("expected a CUDA torch.Tensor")
        if waves.dtype != torch.float32:
            raise ParameterError("Audio data must be float32")
        if waves.dim() == 1:
            waves = waves[None]
        if waves.dim() != 2:
            raise ParameterError("expected (B, n) waveforms")
        if waves.stride(1) != 1:
            waves = waves.contiguous()
        if waves.device.index != self.device:
            raise ParameterError(f"tensor is on cuda:{waves.device.index}, plan on cuda:{self.device}")
        return waves

    def extract_device(self, waves, *, mfcc=True, stats=True, status=True, pooled=False, chroma=False,
                       out=None):
        """(B, n) CUDA float32 -> dict of CUDA tensors; asynchronous on the current stream.

        Keys: ``logmel`` (B, n_mels, T), ``mfcc`` (B, n_mfcc, T), ``stats`` (B, 5, T),
        ``status`` (B,) int32, ``pooled`` (B, 2*n_mels + 2*n_mfcc + 10 [+ 24 with chroma]);
        with ``chroma=True`` also ``chroma`` (B, 12, T) = librosa.feature.chroma_stft and
        ``tuning`` (B,) = librosa.estimate_tuning per clip.  ``out`` may hold preallocated
        tensors under the same keys (plus ``clipmax``).
        """
        import torch

        waves = self._as_cuda_batch(waves)
        B, n = waves.shape
        T = self.num_frames(n)
        dev = waves.device
        out = dict(out) if out else {}
        mfcc = bool(mfcc) and self.n_mfcc > 0

        def buf(key, shape, dtype=torch.float32):
            t = out.get(key)
            if t is None:
                t = torch.empty(shape, dtype=dtype, device=dev)
                out[key] = t
            elif (tuple(t.shape) != tuple(shape) or t.dtype != dtype or not t.is_contiguous()
                  or t.device != dev):
                # a caller-supplied buffer is a promise about where the result lands: never swap it silently
                raise ParameterError(f"out[{key!r}] must be a contiguous {dtype} tensor of shape {tuple(shape)} "
                                     f"on {dev}, got {t.dtype} {tuple(t.shape)} on {t.device}")
            return t

        logmel = buf("logmel", (B, self.n_mels, T))
        mf = buf("mfcc", (B, self.n_mfcc, T)) if mfcc else None
        st = buf("stats", (B, 5, T)) if (stats or pooled) else None
        sta = buf("status", (B,), torch.int32) if status else None
        cm = buf("clipmax", (B,))
        stream = torch.cuda.current_stream(dev).cuda_stream
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        ch = tu = work = None
        wbytes = 0
        if chroma:
            wbytes = int(lib.hlmc_chroma_workspace_bytes(self._plan, B, n))
            if wbytes < 0:
                _check(wbytes)
            ch = buf("chroma", (B, 12, T))
            tu = buf("tuning", (B,))
            work = buf("chroma_work", (wbytes,), torch.uint8)
        _check(lib.hlmc_extract_device_ex(self._plan, ptr(waves), B, n, _row_pitch(waves), ptr(logmel), ptr(mf),
                                          ptr(st), ptr(sta), ptr(cm), ptr(ch), ptr(tu), ptr(work), wbytes,
                                          C.c_void_p(stream)))
        if pooled:
            po = buf("pooled", (B, self.pooled_width(mfcc, bool(chroma))))
            _check(lib.hlmc_pool_device_ex(self._plan, ptr(logmel), ptr(mf), ptr(st), ptr(ch), B, T, ptr(po),
                                           C.c_void_p(stream)))
        return out

    def capture_device(self, waves, *, mfcc=True, stats=True, status=True):
        """Capture ``extract_device`` on this batch as a CUDA graph (small batches are launch-bound).

        Returns a :class:`DeviceGraph`: write new audio into ``graph.waves`` (same shape), call
        ``graph.replay()`` and read ``graph.out`` - the buffers the graph was captured with."""
        import torch

        waves = self._as_cuda_batch(waves)
        if waves.shape[0] == 0:
            raise ParameterError("cannot capture an empty batch")
        out = self.extract_device(waves, mfcc=mfcc, stats=stats, status=status)
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        handle = C.c_void_p()
        B, n = waves.shape
        _check(lib.hlmc_graph_create(self._plan, ptr(waves), B, n, _row_pitch(waves), ptr(out["logmel"]),
                                     ptr(out.get("mfcc")), ptr(out.get("stats")), ptr(out.get("status")),
                                     ptr(out["clipmax"]), C.byref(handle)))
        return DeviceGraph(self, handle, waves, out)

    def extract_pooled_device(self, waves, *, mfcc=True, chroma=False):
        """Only the time-pooled columns of a device-resident batch ([R] extract_all_features /
        extract_flattened_features keep nothing else): ``pooled`` (B, 2*n_mels + 2*n_mfcc + 10 [+ 24])
        and ``status`` (B,).  power_to_db, the DCT and np.mean / np.std over frames run in one kernel;
        the (B, n_mels, T) / (B, n_mfcc, T) arrays are never materialised."""
        import torch

        waves = self._as_cuda_batch(waves)
        B, n = waves.shape
        T = self.num_frames(n)
        dev = waves.device
        mfcc = bool(mfcc) and self.n_mfcc > 0
        new = lambda shape, dt=torch.float32: torch.empty(shape, dtype=dt, device=dev)
        po = new((B, self.pooled_width(mfcc, bool(chroma))))
        st, sta, cm = new((B, 5, T)), new((B,), torch.int32), new((B,))
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        ch = tu = work = None
        wbytes = 0
        if chroma:
            wbytes = int(lib.hlmc_chroma_workspace_bytes(self._plan, B, n))
            if wbytes < 0:
                _check(wbytes)
            ch, tu, work = new((B, 12, T)), new((B,)), new((max(wbytes, 1),), torch.uint8)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(lib.hlmc_extract_pooled_device(self._plan, ptr(waves), B, n, _row_pitch(waves), ptr(po), int(mfcc),
                                              int(bool(chroma)), ptr(st), ptr(sta), ptr(cm), ptr(ch), ptr(tu),
                                              ptr(work), wbytes, C.c_void_p(stream)))
        return {"pooled": po, "status": sta}

    def melspectrogram_device(self, waves, *, stats=False):
        import torch

        waves = self._as_cuda_batch(waves)
        B, n = waves.shape
        T = self.num_frames(n)
        mel = torch.empty((B, self.n_mels, T), dtype=torch.float32, device=waves.device)
        st = torch.empty((B, 5, T), dtype=torch.float32, device=waves.device) if stats else None
        stream = torch.cuda.current_stream(waves.device).cuda_stream
        _check(lib.hlmc_melspectrogram_device(
            self._plan, C.c_void_p(waves.data_ptr()), B, n, _row_pitch(waves), C.c_void_p(mel.data_ptr()),
            C.c_void_p(st.data_ptr()) if st is not None else None, None, C.c_void_p(stream)))
        return (mel, st) if stats else mel

    def stats_device(self, waves):
        """Only the five spectral / temporal statistics, (B, 5, T)."""
        import torch

        waves = self._as_cuda_batch(waves)
        B, n = waves.shape
        T = self.num_frames(n)
        st = torch.empty((B, 5, T), dtype=torch.float32, device=waves.device)
        stream = torch.cuda.current_stream(waves.device).cuda_stream
        _check(lib.hlmc_melspectrogram_device(
            self._plan, C.c_void_p(waves.data_ptr()), B, n, _row_pitch(waves), None,
            C.c_void_p(st.data_ptr()), None, C.c_void_p(stream)))
        return st

    def stft_device(self, waves):
        import torch

        waves = self._as_cuda_batch(waves)
        B, n = waves.shape
        T = self.num_frames(n)
        spec = torch.empty((B, self.n_bins, T, 2), dtype=torch.float32, device=waves.device)
        stream = torch.cuda.current_stream(waves.device).cuda_stream
        _check(lib.hlmc_stft_device(self._plan, C.c_void_p(waves.data_ptr()), B, n, _row_pitch(waves),
                                    C.c_void_p(spec.data_ptr()), C.c_void_p(stream)))
        return torch.view_as_complex(spec)

    # -- host path (the reference-facing call) ---------------------------------
    def extract_host(self, waves, *, logmel=True, mfcc=True, stats=True, status=True, pooled=False,
                     chroma=False, chunk_clips=0, n_streams=3, out=None, pad_to=None, fixed_frames=None,
                     sr_in=None, wave_out=False, valid_frames=None):
        """(B, n) host float32 -- or int16 PCM -- (numpy or CPU torch, ideally pinned) -> dict of numpy arrays.

        H2D copies, kernels and D2H copies are overlapped inside the C library.  The front end of the
        scripts' ``load_audio_file`` ([R] src/1_preprocessing.py:137-153) runs on the device:

        * int16 input is converted as librosa.load does for PCM16 files (x / 32768);
        * a 3-D ``(B, n, channels)`` input holds interleaved frames and is averaged to mono (``librosa.to_mono``);
        * ``sr_in`` different from the plan's ``sr`` resamples with ``librosa.resample(res_type="polyphase")``
          (= ``scipy.signal.resample_poly``; librosa.load's default ``soxr_hq`` is a different low-pass);
        * ``pad_to`` right zero-pads every clip to that many samples (at the plan's rate), the scripts'
          ``np.pad`` to ``sample_rate * duration``.

        ``valid_frames`` (B,) gives every clip its own frame count (files shorter than ``duration``): frames
        past it are ignored and the clip is zero-padded after ITS resampled length.

        ``fixed_frames`` adds ``fixed_logmel`` (B, n_mels, fixed_frames): the log-mel image cropped or padded
        with its minimum ([R] src/1_preprocessing_advanced.py:108-112).  ``wave_out=True`` also returns
        ``wave`` (B, n_total), the clips as the features saw them.  Entries of ``out`` must have exactly the
        shape and dtype the call produces (``ParameterError`` otherwise).
        """
        tensor_in = None
        try:
            import torch

            if isinstance(waves, torch.Tensor):
                if waves.is_cuda:
                    raise TypeError("extract_host expects host memory; use extract_device")
                tensor_in = waves
                waves = waves.numpy()
        except ImportError:  # pragma: no cover
            pass
        waves = np.asarray(waves)
        pcm16 = waves.dtype == np.int16
        if not pcm16:
            if not np.issubdtype(waves.dtype, np.floating):
                raise ParameterError("Audio data must be floating-point")
            if waves.dtype != np.float32:
                waves = waves.astype(np.float32)
        esz = waves.dtype.itemsize
        if waves.ndim == 1:
            waves = waves[None]
        channels = 1
        if waves.ndim == 3:                       # (B, n, channels): frames interleaved as in a WAV file
            channels = int(waves.shape[2])
            if channels < 1:
                raise ParameterError("expected at least one channel")
            if not waves.flags.c_contiguous:
                waves = np.ascontiguousarray(waves)
        elif waves.ndim != 2:
            raise ParameterError("expected (B, n) or (B, n, channels) waveforms")
        fsz = esz * channels
        if channels == 1:
            waves = waves.reshape(waves.shape[0], waves.shape[1])
            if waves.strides[1] != esz or (waves.shape[0] > 1 and (waves.strides[0] % esz or waves.strides[0] < esz * waves.shape[1])):
                waves = np.ascontiguousarray(waves)
        B, n = waves.shape[0], waves.shape[1]
        sr_in = int(sr_in) if sr_in else 0
        if sr_in == self.params.sr:
            sr_in = 0
        n_res = int(lib.hlmc_resampled_length(n, sr_in, self.params.sr)) if sr_in else n
        n_total = int(pad_to) if pad_to else n_res
        if n_total < n_res:
            raise ParameterError("pad_to is shorter than the clips")
        T = self.num_frames(n_total)
        mfcc = bool(mfcc) and self.n_mfcc > 0
        out = dict(out) if out else {}

        def buf(key, shape, dtype=np.float32):
            a = out.get(key)
            if a is None:
                a = np.empty(shape, dtype=dtype)
                out[key] = a
            elif not (isinstance(a, np.ndarray) and a.shape == tuple(shape) and a.dtype == dtype
                      and a.flags.c_contiguous and a.flags.writeable):
                raise ParameterError(f"out[{key!r}] must be a writable C-contiguous {np.dtype(dtype).name} array of "
                                     f"shape {tuple(shape)}, got {getattr(a, 'dtype', type(a))} {getattr(a, 'shape', '')}")
            return a

        lm = buf("logmel", (B, self.n_mels, T)) if logmel else None
        mf = buf("mfcc", (B, self.n_mfcc, T)) if mfcc else None
        st = buf("stats", (B, 5, T)) if stats else None
        sta = buf("status", (B,), np.int32) if status else None
        po = buf("pooled", (B, self.pooled_width(self.n_mfcc > 0, bool(chroma)))) if pooled else None
        ch = buf("chroma", (B, 12, T)) if (chroma and chroma != "pooled") else None     # "pooled": columns only
        tu = buf("tuning", (B,)) if (chroma and chroma != "pooled") else None
        fx = buf("fixed_logmel", (B, self.n_mels, int(fixed_frames))) if fixed_frames else None
        wo = buf("wave", (B, n_total)) if wave_out else None
        vf = None
        if valid_frames is not None:
            vf = np.ascontiguousarray(valid_frames, dtype=np.int64)
            if vf.shape != (B,) or (vf < 0).any() or (vf > n).any():
                raise ParameterError("valid_frames must be (B,) counts in [0, n]")
        ptr = lambda a: a.ctypes.data if a is not None else None
        io = HlmcHostIo(wave=ptr(waves), sample_format=1 if pcm16 else 0, B=B, n_valid=n,
                        pitch=n if B <= 1 else waves.strides[0] // fsz, n_total=n_total, logmel=ptr(lm),
                        mfcc=ptr(mf), stats=ptr(st), chroma=ptr(ch), tuning=ptr(tu), pooled=ptr(po),
                        status=ptr(sta), pooled_with_chroma=int(bool(chroma)), chunk_clips=int(chunk_clips),
                        n_streams=int(n_streams), channels=channels, sr_in=sr_in, reserved0=0,
                        fixed_logmel=ptr(fx), fixed_frames=int(fixed_frames or 0), wave_out=ptr(wo),
                        valid_frames=ptr(vf))
        with self._lock:
            _check(lib.hlmc_extract_host_io(self._plan, C.byref(io)))
        del tensor_in
        return out

    def load_frontend_device(self, raw, *, sr_in=None, pad_to=None, valid_frames=None):
        """librosa.load's arithmetic for a device-resident batch: (B, n) or (B, n, channels) CUDA int16 /
        float32 -> (B, n_total) float32 at the plan's rate (mono mix, polyphase resampling, zero pad)."""
        import torch

        if not (isinstance(raw, torch.Tensor) and raw.is_cuda):
            raise TypeError("expected a CUDA torch.Tensor")
        if raw.dtype not in (torch.int16, torch.float32):
            raise ParameterError("Audio data must be int16 PCM or float32")
        if raw.dim() == 2:
            raw = raw[:, :, None]
        if raw.dim() != 3:
            raise ParameterError("expected (B, n) or (B, n, channels)")
        raw = raw.contiguous()
        B, n, channels = raw.shape
        sr_in = int(sr_in) if sr_in else 0
        n_res = int(lib.hlmc_resampled_length(n, sr_in, self.params.sr)) if sr_in else n
        n_total = int(pad_to) if pad_to else n_res
        if n_total < n_res:
            raise ParameterError("pad_to is shorter than the clips")
        pitch = (n_total + 3) & ~3
        store = torch.empty((B, pitch), dtype=torch.float32, device=raw.device)
        stream = torch.cuda.current_stream(raw.device).cuda_stream
        vf = None
        if valid_frames is not None:
            vf = torch.as_tensor(valid_frames, dtype=torch.int64).to(raw.device).contiguous()
        _check(lib.hlmc_load_frontend_device(self._plan, C.c_void_p(raw.data_ptr()), 1 if raw.dtype == torch.int16 else 0,
                                             channels, B, n, n, sr_in, C.c_void_p(store.data_ptr()), pitch, n_total,
                                             C.c_void_p(vf.data_ptr()) if vf is not None else None,
                                             C.c_void_p(stream)))
        return store[:, :n_total]

    def last_transfer_bytes(self):
        h2d, d2h = C.c_int64(0), C.c_int64(0)
        lib.hlmc_last_transfer_bytes(self._plan, C.byref(h2d), C.byref(d2h))
        return int(h2d.value), int(d2h.value)

    def extract(self, waves, **kw):
        """Dispatch on where ``waves`` lives."""
        try:
            import torch

            if isinstance(waves, torch.Tensor) and waves.is_cuda:
                return self.extract_device(waves, **kw)
        except ImportError:  # pragma: no cover
            pass
        return self.extract_host(waves, **kw)


_CACHE: "OrderedDict" = None
_CACHE_LOCK = threading.Lock()
_CACHE_MAX = 32


def get_extractor(**kw) -> FeatureExtractor:
    """Process-wide plan cache keyed by the parameter set (plans are cheap but not free).

    Callable and array windows are keyed by the window they resolve to (two lambdas never share a plan);
    the cache holds at most ``_CACHE_MAX`` plans and closes the least recently used one beyond that."""
    global _CACHE
    from collections import OrderedDict

    def freeze(k, v):
        if k == "window" and not isinstance(v, str):
            wl = kw.get("win_length") or kw.get("n_fft", 2048)
            w = _resolve_window(v, int(wl))
            return ("win", None if w is None else w.tobytes())
        if isinstance(v, np.ndarray):
            return ("nd", v.shape, v.dtype.str, v.tobytes())
        if callable(v):
            if v in (np.max, np.amax, max):
                return ("fn", "max")
            return ("fn", id(v))
        if isinstance(v, list):
            return tuple(v)
        return v

    key = tuple(sorted((k, freeze(k, v)) for k, v in kw.items()))
    with _CACHE_LOCK:
        if _CACHE is None:
            _CACHE = OrderedDict()
        ex = _CACHE.get(key)
        if ex is None:
            ex = FeatureExtractor(**kw)
            _CACHE[key] = ex
            while len(_CACHE) > _CACHE_MAX:
                _k, old = _CACHE.popitem(last=False)
                old.close()
        else:
            _CACHE.move_to_end(key)
        return ex


def launch_count() -> int:
    return int(lib.hlmc_launch_count())


def measure_fp32_peak(device=0) -> float:
    v = C.c_double(0.0)
    _check(lib.hlmc_measure_fp32_peak(int(device), C.byref(v)))
    return float(v.value)
