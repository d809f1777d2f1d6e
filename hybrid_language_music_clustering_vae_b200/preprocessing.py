"""The reference scripts' feature helpers and on-disk layout, driven by the CUDA path.

Function names, arguments and return shapes follow
``src/1_preprocessing.py`` (``processed_data1/``) and
``src/1_preprocessing_advanced.py`` (``processed_data2/``); the batched
functions replace the scripts' per-file loop / joblib map with one call.

Chroma: the saved vectors carry 24 ``chroma_stft`` columns
([R] 1_preprocessing.py:126-127, _advanced.py:153-154).  ``chroma="device"``
(default) computes them on the GPU (librosa.feature.chroma_stft including the
per-clip tuning estimate); ``"zeros"`` / ``"nan"`` skip that
second STFT pass and fill the columns (logged), and an explicit (B, 24) array
is taken as given.  There is never a hidden CPU path.
"""
from __future__ import annotations

import logging
import os
import pickle
import wave as _wave

import numpy as np

from .core import FeatureExtractor, ParameterError, get_extractor, STAT_NAMES

log = logging.getLogger(__name__)

# [R] src/1_preprocessing.py:21-29
BASIC_CONFIG = {
    "sample_rate": 22050, "duration": 30, "n_mels": 128, "n_fft": 2048, "hop_length": 512,
    "n_mfcc": 40, "max_samples_per_class": 160,
}
# [R] src/1_preprocessing_advanced.py:28-37
ADV_CONFIG = {
    "sample_rate": 22050, "duration": 30, "n_mels": 128, "n_fft": 2048, "hop_length": 512,
    "fixed_time_steps": 1024, "max_samples_per_class": 200, "lyrics_max_features": 768,
}
N_CHROMA_COLS = 24


def _basic_extractor(cfg=BASIC_CONFIG, device=0) -> FeatureExtractor:
    # melspectrogram(n_mels, n_fft, hop) + power_to_db(ref=np.max); mfcc(n_mfcc, n_fft, hop)
    return get_extractor(sr=cfg["sample_rate"], n_fft=cfg["n_fft"], hop_length=cfg["hop_length"],
                         n_mels=cfg["n_mels"], n_mfcc=cfg.get("n_mfcc", 0), ref=np.max, device=device)


def _stats_extractor(cfg=BASIC_CONFIG, device=0) -> FeatureExtractor:
    # the five statistics are called WITHOUT n_fft ([R] 1_preprocessing.py:75-83, _advanced.py:133-137):
    # librosa's default n_fft = frame_length = 2048 whatever CONFIG['n_fft'] says; only hop_length is forwarded
    return get_extractor(sr=cfg["sample_rate"], n_fft=2048, hop_length=cfg["hop_length"], n_mfcc=0, device=device)


def _front_kw(kw):
    return {k: kw[k] for k in ("sr_in", "valid_frames", "pad_to") if kw.get(k) is not None}


def _chroma_block(chroma, B):
    if isinstance(chroma, str):
        if chroma == "zeros":
            log.warning("chroma_stft columns are filled with zeros (not computed by this path)")
            return np.zeros((B, N_CHROMA_COLS))
        if chroma == "nan":
            log.warning("chroma_stft columns are filled with NaN (not computed by this path)")
            return np.full((B, N_CHROMA_COLS), np.nan)
        raise ValueError(f"chroma policy {chroma!r}")
    c = np.asarray(chroma, dtype=np.float64)
    if c.shape != (B, N_CHROMA_COLS):
        raise ValueError(f"chroma block must be ({B}, {N_CHROMA_COLS})")
    return c


def _raise_failed(status):
    if np.any(status):
        raise ParameterError("Audio buffer is not finite everywhere")


def _pooled_rows(waves, cfg, n_mfcc, chroma, device, fixed_frames=None, **front):
    """One pass of the host pipeline -> (pooled float64 (B, W), status, fixed image or None).

    W = 2*n_mels + 2*n_mfcc + 10 + 24.  When CONFIG['n_fft'] is not 2048 the ten statistic columns come
    from a second plan with librosa's default n_fft = 2048, as in the reference."""
    ex = get_extractor(sr=cfg["sample_rate"], n_fft=cfg["n_fft"], hop_length=cfg["hop_length"],
                       n_mels=cfg["n_mels"], n_mfcc=n_mfcc, ref=np.max, device=device)
    on_dev = isinstance(chroma, str) and chroma == "device"
    r = ex.extract_host(waves, logmel=False, mfcc=False, stats=False, pooled=True,
                        chroma="pooled" if on_dev else False, fixed_frames=fixed_frames, **front)
    status = r["status"]
    feats = r["pooled"].astype(np.float64)
    B = feats.shape[0]
    if cfg["n_fft"] != 2048:
        st = _stats_extractor(cfg, device).extract_host(waves, logmel=False, mfcc=False, **front)
        status = status | st["status"]
        s64 = st["stats"].astype(np.float64)
        o = 2 * ex.n_mels + 2 * ex.n_mfcc
        feats[:, o:o + 10:2] = s64.mean(axis=2)
        feats[:, o + 1:o + 10:2] = s64.std(axis=2)
    if not on_dev:
        feats = np.concatenate([feats, _chroma_block(chroma, B)], axis=1)
    feats[status != 0] = np.nan
    return feats, status, r.get("fixed_logmel")


# ---------------------------------------------------------------------------
# src/1_preprocessing.py
# ---------------------------------------------------------------------------
def extract_mel_spectrogram(audio, sr, cfg=BASIC_CONFIG):
    """[R] 1_preprocessing.py:48-58 -> (n_mels, T) float32 log-mel, ref=max."""
    ex = _basic_extractor(dict(cfg, sample_rate=sr))
    r = ex.extract_host(np.asarray(audio)[None], mfcc=False, stats=False)
    _raise_failed(r["status"])
    return r["logmel"][0]


def extract_mfcc(audio, sr, cfg=BASIC_CONFIG):
    """[R] 1_preprocessing.py:61-70 -> (n_mfcc, T) float32."""
    ex = _basic_extractor(dict(cfg, sample_rate=sr))
    r = ex.extract_host(np.asarray(audio)[None], logmel=False, stats=False)
    _raise_failed(r["status"])
    return r["mfcc"][0]


def extract_spectral_features(audio, sr, cfg=BASIC_CONFIG):
    """[R] 1_preprocessing.py:73-91 -> dict of (1, T) arrays (float64; rms float32)."""
    ex = _stats_extractor(dict(cfg, sample_rate=sr))
    r = ex.extract_host(np.asarray(audio)[None], logmel=False, mfcc=False)
    _raise_failed(r["status"])
    st = r["stats"][0]
    return {name: (st[i:i + 1].astype(np.float64) if name != "rms" else st[i:i + 1].copy())
            for i, name in enumerate(STAT_NAMES)}


def extract_all_features_batch(waves, sr=22050, cfg=BASIC_CONFIG, chroma="device", device=0,
                               return_status=False, **front):
    """Batched [R] 1_preprocessing.py:105-129: (B, n) -> (B, 370) float64.

    Columns: mel mean/std (256) | MFCC mean/std (80) | 5 x (mean, std) | chroma (24).
    Rows whose clip was non-finite are returned as NaN and flagged in ``status``
    (the script skips such files, [R] 1_preprocessing.py:248-251).  ``front``: ``sr_in``,
    ``valid_frames``, ``pad_to`` of ``FeatureExtractor.extract_host`` (raw PCM / other-rate input)."""
    feats, status, _ = _pooled_rows(waves, dict(cfg, sample_rate=sr), cfg.get("n_mfcc", 40), chroma, device,
                                    **_front_kw(front))
    return (feats, status) if return_status else feats


def extract_chroma_features(audio, sr, cfg=BASIC_CONFIG):
    """[R] 1_preprocessing.py:94-102 -> (12, T) float32 chroma_stft."""
    from .api import feature

    return feature.chroma_stft(y=np.asarray(audio), sr=sr, n_fft=cfg["n_fft"], hop_length=cfg["hop_length"])


def extract_all_features(audio, sr, cfg=BASIC_CONFIG, chroma="device"):
    """[R] 1_preprocessing.py:105-129 -> (370,) float64."""
    f, status = extract_all_features_batch(np.asarray(audio)[None], sr, cfg, chroma, return_status=True)
    _raise_failed(status)
    return f[0]


# ---------------------------------------------------------------------------
# src/1_preprocessing_advanced.py
# ---------------------------------------------------------------------------
def process_batch_advanced(waves, sr=22050, cfg=ADV_CONFIG, chroma="device", device=0, **front):
    """Batched [R] _advanced.py:97-156 (extract_mel_spectrogram + extract_flattened_features).

    (B, n) host clips -> (mel (B, n_mels, fixed_time_steps) f32, flat (B, 290) f64, status (B,)), through the
    library's 3-stream host pipeline (H2D | kernels | D2H overlapped); the crop / pad-with-min of the image
    runs on the device."""
    feats, status, mel = _pooled_rows(waves, dict(cfg, sample_rate=sr), 0, chroma, device,
                                      fixed_frames=int(cfg["fixed_time_steps"]), **_front_kw(front))
    return mel, feats, status


def extract_mel_spectrogram_fixed(audio, sr, cfg=ADV_CONFIG):
    """[R] _advanced.py:97-114 -> (n_mels, fixed_time_steps) float32."""
    ex = get_extractor(sr=sr, n_fft=cfg["n_fft"], hop_length=cfg["hop_length"], n_mels=cfg["n_mels"], n_mfcc=0,
                       ref=np.max, device=0)
    r = ex.extract_host(np.asarray(audio)[None], logmel=False, mfcc=False, stats=False,
                        fixed_frames=int(cfg["fixed_time_steps"]))
    _raise_failed(r["status"])
    return r["fixed_logmel"][0]


def extract_flattened_features(audio, sr, cfg=ADV_CONFIG, chroma="device"):
    """[R] _advanced.py:120-156 -> (290,) float64."""
    f, status, _ = _pooled_rows(np.asarray(audio)[None], dict(cfg, sample_rate=sr), 0, chroma, 0)
    _raise_failed(status)
    return f[0]


# ---------------------------------------------------------------------------
# load_audio_file / process_single_file ([R] 1_preprocessing.py:137-153; _advanced.py:79-94, 158-183).
# The WAV container is parsed on the host (stdlib `wave`: header + raw PCM16 frames, no arithmetic);
# everything librosa.load computes -- int16 -> float32, mono mix, resampling, crop to `duration`, zero pad --
# runs on the device.  Resampling is librosa's res_type="polyphase" (see FeatureExtractor.extract_host).
# ---------------------------------------------------------------------------
def read_wav_pcm16(file_path, duration=None):
    """-> (frames int16 (n, channels), native sample rate); at most ``duration`` seconds, as librosa.load reads."""
    with _wave.open(str(file_path), "rb") as w:
        if w.getsampwidth() != 2 or w.getcomptype() != "NONE":
            raise ValueError(f"{file_path}: only uncompressed 16-bit PCM WAV files are decoded here")
        sr, ch, n = w.getframerate(), w.getnchannels(), w.getnframes()
        if duration is not None:
            n = min(n, int(duration * sr))
        raw = w.readframes(n)
    a = np.frombuffer(raw, dtype="<i2")
    return a.reshape(-1, ch), sr


def load_audio_file(file_path, cfg=BASIC_CONFIG, device=0):
    """[R] 1_preprocessing.py:137-153 -> (audio float32 (sample_rate*duration,), sr), or (None, None)."""
    import torch

    try:
        frames, sr_native = read_wav_pcm16(file_path, cfg["duration"])
        ex = _stats_extractor(cfg, device)
        expected = cfg["sample_rate"] * cfg["duration"]
        n_res = int(np.ceil(len(frames) * cfg["sample_rate"] / sr_native))
        y = ex.load_frontend_device(torch.from_numpy(frames.copy())[None].to(torch.device("cuda", device)),
                                    sr_in=sr_native, pad_to=max(expected, n_res))
        return y[0].cpu().numpy(), cfg["sample_rate"]
    except Exception as e:
        print(f"Error loading {file_path}: {e}")
        return None, None


def _decode_groups(file_infos, cfg):
    """Parse the containers and group the clips by (native rate, channels).  -> (groups, load_failed indices)"""
    groups, failed = {}, []
    for i, info in enumerate(file_infos):
        try:
            frames, sr_native = read_wav_pcm16(info["path"], cfg["duration"])
            if len(frames) == 0:
                raise ValueError("empty file")
            groups.setdefault((sr_native, frames.shape[1]), []).append((i, frames))
        except Exception as e:
            print(f"Error loading {info['path']}: {e}")
            failed.append(i)
    return groups, failed


def _stack_group(items, sr_native, cfg):
    n_max = max(len(f) for _i, f in items)
    raw = np.zeros((len(items), n_max, items[0][1].shape[1]), np.int16)
    valid = np.empty((len(items),), np.int64)
    for k, (_i, f) in enumerate(items):
        raw[k, :len(f)] = f
        valid[k] = len(f)
    expected = cfg["sample_rate"] * cfg["duration"]
    n_res = int(np.ceil(n_max * cfg["sample_rate"] / sr_native))
    return raw, valid, max(expected, n_res)


def process_files_basic(file_infos, cfg=BASIC_CONFIG, chroma="device", device=0):
    """The basic script's per-file loop ([R] 1_preprocessing.py:223-258: load_audio_file + extract_all_features)
    as batched device calls.  -> (features (N, 370) float64, ok (N,) bool, error text per file)."""
    n = len(file_infos)
    feats = np.full((n, 2 * cfg["n_mels"] + 2 * cfg.get("n_mfcc", 40) + 34), np.nan)
    ok = np.zeros((n,), bool)
    errors = [""] * n
    groups, failed = _decode_groups(file_infos, cfg)
    for i in failed:
        errors[i] = "Failed to load"
    for (sr_native, _ch), items in groups.items():
        raw, valid, pad_to = _stack_group(items, sr_native, cfg)
        try:
            f, status = extract_all_features_batch(raw, cfg["sample_rate"], cfg, chroma, device, return_status=True,
                                                   sr_in=sr_native, valid_frames=valid, pad_to=pad_to)
        except Exception as e:
            for i, _f in items:
                errors[i] = str(e)
            continue
        for k, (i, _f) in enumerate(items):
            if status[k]:
                errors[i] = "Audio buffer is not finite everywhere"
            else:
                feats[i], ok[i] = f[k], True
    return feats, ok, errors


def process_files_advanced(file_infos, cfg=ADV_CONFIG, chroma="device", device=0):
    """The advanced script's ``Parallel(...)(delayed(process_single_file)(f) ...)`` map ([R] _advanced.py:286-288)
    as batched device calls: files are decoded on the host, grouped by (rate, channels) and run through the
    host pipeline with per-clip lengths.  Returns the list of per-file dicts ``process_single_file`` returns."""
    results = [None] * len(file_infos)
    groups, failed = _decode_groups(file_infos, cfg)
    for i in failed:
        results[i] = {"status": "failed", "path": file_infos[i]["path"], "error": "Load failed"}
    for (sr_native, _ch), items in groups.items():
        raw, valid, pad_to = _stack_group(items, sr_native, cfg)
        try:
            mel, flat, status = process_batch_advanced(raw, cfg["sample_rate"], cfg, chroma, device, sr_in=sr_native,
                                                       valid_frames=valid, pad_to=pad_to)
        except Exception as e:
            for i, _f in items:
                results[i] = {"status": "failed", "path": file_infos[i]["path"], "error": str(e)}
            continue
        for k, (i, _f) in enumerate(items):
            info = file_infos[i]
            if status[k]:
                results[i] = {"status": "failed", "path": info["path"], "error": "Audio buffer is not finite everywhere"}
            else:
                results[i] = {"status": "success", "mel_spec": mel[k], "flat_feat": flat[k], "genre": info["genre"],
                              "lyrics": info["lyrics"], "language": info["language"], "filename": info["filename"],
                              "file_id": info["file_id"]}
    return results


def process_single_file(file_info, cfg=ADV_CONFIG, chroma="device", device=0):
    """[R] _advanced.py:158-183: one file -> the worker's result dict (same keys, same failure records)."""
    return process_files_advanced([file_info], cfg, chroma, device)[0]


# ---------------------------------------------------------------------------
# on-disk layout (SURVEY.md 8a "layout")
# ---------------------------------------------------------------------------
def _normalise_tabular(features, device):
    """inf -> nan, SimpleImputer(strategy="mean"), StandardScaler, as [R] 1_preprocessing.py:303-311.
    device=None: sklearn on the host (the reference's own code path); an ordinal: the CUDA kernels, with
    sklearn objects rebuilt around the device statistics.  -> (normalized, imputer, scaler)."""
    from sklearn.impute import SimpleImputer
    from sklearn.preprocessing import StandardScaler

    if device is None:
        features_clean = np.where(np.isinf(features), np.nan, features)
        imputer = SimpleImputer(strategy="mean")
        features_imputed = imputer.fit_transform(features_clean)
        scaler = StandardScaler()
        return scaler.fit_transform(features_imputed), imputer, scaler
    import torch
    from .scaler import fit_transform_tabular_device

    x = torch.from_numpy(np.ascontiguousarray(features, dtype=np.float64)).to(torch.device("cuda", device))
    _imp, scaled, imputer, scaler = fit_transform_tabular_device(x)
    return scaled.cpu().numpy(), imputer, scaler


def save_processed_data1(out_dir, features, labels, metadata_df=None, config=BASIC_CONFIG, device=None):
    """[R] 1_preprocessing.py:295-343, same files and contents: ``features_raw.npy`` is the matrix AS EXTRACTED
    (NaN / inf kept), ``features_normalized.npy`` = StandardScaler(SimpleImputer(inf -> nan)), and the pickles
    are that default ``SimpleImputer(strategy="mean")`` and ``StandardScaler``."""
    os.makedirs(out_dir, exist_ok=True)
    features = np.asarray(features)
    features_normalized, imputer, scaler = _normalise_tabular(features, device)
    np.save(os.path.join(out_dir, "features_raw.npy"), features)
    np.save(os.path.join(out_dir, "features_normalized.npy"), features_normalized)
    np.save(os.path.join(out_dir, "labels.npy"), np.array(labels))
    if metadata_df is not None:
        metadata_df.to_csv(os.path.join(out_dir, "metadata.csv"), index=False)
    for name, obj in (("scaler.pkl", scaler), ("imputer.pkl", imputer), ("config.pkl", dict(config))):
        with open(os.path.join(out_dir, name), "wb") as f:
            pickle.dump(obj, f)
    return features, features_normalized


def save_processed_data2(out_dir, mel_spectrograms, flat_features, labels, lyrics_embeddings=None,
                         metadata_df=None, config=ADV_CONFIG, device_scaler=None):
    """[R] _advanced.py:376-421, same files and contents: raw arrays as extracted, the (N, 131072) mel
    StandardScaler, and inf -> nan / default SimpleImputer / StandardScaler for the 290-vectors.

    ``device_scaler``: CUDA device ordinal to run both normalisations on the GPU; None keeps
    them in sklearn on the host.  Either way the pickles are sklearn objects."""
    from sklearn.preprocessing import StandardScaler

    os.makedirs(out_dir, exist_ok=True)
    mel = np.asarray(mel_spectrograms)
    n, h, w = mel.shape
    if device_scaler is None:
        mel_scaler = StandardScaler()
        mel_norm = mel_scaler.fit_transform(mel.reshape(n, -1)).reshape(n, h, w)
    else:
        import torch
        from .scaler import fit_transform_device

        xd = torch.from_numpy(np.ascontiguousarray(mel.reshape(n, -1), dtype=np.float32)).to(
            torch.device("cuda", device_scaler))
        yd, mel_scaler = fit_transform_device(xd, inplace=True)
        mel_norm = yd.cpu().numpy().reshape(n, h, w)
    flat = np.asarray(flat_features)
    flat_norm, imputer, flat_scaler = _normalise_tabular(flat, device_scaler)
    np.save(os.path.join(out_dir, "mel_spectrograms_raw.npy"), mel)
    np.save(os.path.join(out_dir, "mel_spectrograms_normalized.npy"), mel_norm)
    np.save(os.path.join(out_dir, "features_raw.npy"), flat)
    np.save(os.path.join(out_dir, "features_normalized.npy"), flat_norm)
    if lyrics_embeddings is not None:
        np.save(os.path.join(out_dir, "lyrics_embeddings.npy"), np.asarray(lyrics_embeddings))
    np.save(os.path.join(out_dir, "labels.npy"), np.asarray(labels))
    if metadata_df is not None:
        metadata_df.to_csv(os.path.join(out_dir, "metadata.csv"), index=False)
    for name, obj in (("mel_scaler.pkl", mel_scaler), ("flat_scaler.pkl", flat_scaler),
                      ("imputer.pkl", imputer), ("config.pkl", dict(config))):
        with open(os.path.join(out_dir, name), "wb") as f:
            pickle.dump(obj, f)
    return mel_norm, flat_norm
