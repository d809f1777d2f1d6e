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
    # melspectrogram(n_mels, n_fft, hop) + power_to_db(ref=np.max); mfcc(n_mfcc, n_fft, hop);
    # the spectral statistics use librosa's default n_fft=2048 (the scripts do not forward it)
    return get_extractor(sr=cfg["sample_rate"], n_fft=cfg["n_fft"], hop_length=cfg["hop_length"],
                         n_mels=cfg["n_mels"], n_mfcc=cfg.get("n_mfcc", 0), ref=np.max, device=device)


def _pooled_on_device(ex, waves, with_chroma, device, chunk_clips):
    """Chunked device-resident extraction of the pooled feature rows (+ status)."""
    import torch

    waves = np.ascontiguousarray(waves, dtype=np.float32)
    B = waves.shape[0]
    width = ex.pooled_width(ex.n_mfcc > 0, with_chroma)
    pooled = np.empty((B, width), np.float32)
    status = np.empty((B,), np.int32)
    dev = torch.device("cuda", device)
    for lo in range(0, B, chunk_clips):
        hi = min(B, lo + chunk_clips)
        r = ex.extract_device(torch.from_numpy(waves[lo:hi]).to(dev), mfcc=ex.n_mfcc > 0, stats=True,
                              pooled=True, chroma=with_chroma)
        pooled[lo:hi] = r["pooled"].cpu().numpy()
        status[lo:hi] = r["status"].cpu().numpy()
    return pooled, status


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
    ex = _basic_extractor(dict(cfg, sample_rate=sr, n_fft=2048))
    r = ex.extract_host(np.asarray(audio)[None], logmel=False, mfcc=False)
    _raise_failed(r["status"])
    st = r["stats"][0]
    return {name: (st[i:i + 1].astype(np.float64) if name != "rms" else st[i:i + 1].copy())
            for i, name in enumerate(STAT_NAMES)}


def extract_all_features_batch(waves, sr=22050, cfg=BASIC_CONFIG, chroma="device", device=0,
                               return_status=False, chunk_clips=256):
    """Batched [R] 1_preprocessing.py:105-129: (B, n) -> (B, 370) float64.

    Columns: mel mean/std (256) | MFCC mean/std (80) | 5 x (mean, std) | chroma (24).
    Rows whose clip was non-finite are returned as NaN and flagged in ``status``
    (the script skips such files, [R] 1_preprocessing.py:248-251).
    """
    ex = _basic_extractor(dict(cfg, sample_rate=sr), device=device)
    waves = np.asarray(waves)
    if isinstance(chroma, str) and chroma == "device":
        r = ex.extract_host(waves, logmel=False, mfcc=False, stats=False, pooled=True, chroma="pooled")
        status = r["status"]
        feats = r["pooled"].astype(np.float64)
    else:
        r = ex.extract_host(waves, logmel=False, mfcc=False, stats=False, pooled=True)
        status = r["status"]
        feats = np.concatenate([r["pooled"].astype(np.float64), _chroma_block(chroma, waves.shape[0])], axis=1)
    feats[status != 0] = np.nan
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
def process_batch_advanced(waves, sr=22050, cfg=ADV_CONFIG, chroma="device", device=0, chunk_clips=64):
    """Batched [R] _advanced.py:97-156 (extract_mel_spectrogram + extract_flattened_features).

    (B, n) host float32 -> (mel (B, n_mels, fixed_time_steps) f32, flat (B, 290) f64, status (B,)).
    """
    import ctypes as C
    import torch
    from ._lib import lib
    from .core import _check

    ex = get_extractor(sr=sr, n_fft=cfg["n_fft"], hop_length=cfg["hop_length"], n_mels=cfg["n_mels"],
                       n_mfcc=0, ref=np.max, device=device)
    waves = np.ascontiguousarray(waves, dtype=np.float32)
    B, n = waves.shape
    T = ex.num_frames(n)
    fixed = int(cfg["fixed_time_steps"])
    mel = np.empty((B, ex.n_mels, fixed), np.float32)
    on_dev = isinstance(chroma, str) and chroma == "device"
    flat = np.empty((B, 2 * ex.n_mels + 10 + (24 if on_dev else 0)), np.float32)
    status = np.empty((B,), np.int32)
    dev = torch.device("cuda", device)
    for lo in range(0, B, chunk_clips):
        hi = min(B, lo + chunk_clips)
        w = torch.from_numpy(waves[lo:hi]).to(dev, non_blocking=True)
        r = ex.extract_device(w, mfcc=False, stats=True, pooled=True, chroma=on_dev)
        fx = torch.empty((hi - lo, ex.n_mels, fixed), dtype=torch.float32, device=dev)
        _check(lib.hlmc_fix_frames_device(C.c_void_p(r["logmel"].data_ptr()), C.c_void_p(fx.data_ptr()),
                                          hi - lo, ex.n_mels, T, fixed, device,
                                          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        mel[lo:hi] = fx.cpu().numpy()
        flat[lo:hi] = r["pooled"].cpu().numpy()
        status[lo:hi] = r["status"].cpu().numpy()
    feats = flat.astype(np.float64) if on_dev else np.concatenate(
        [flat.astype(np.float64), _chroma_block(chroma, B)], axis=1)
    feats[status != 0] = np.nan
    return mel, feats, status


def extract_mel_spectrogram_fixed(audio, sr, cfg=ADV_CONFIG):
    """[R] _advanced.py:97-114 -> (n_mels, fixed_time_steps) float32."""
    mel, _f, status = process_batch_advanced(np.asarray(audio)[None], sr, cfg)
    _raise_failed(status)
    return mel[0]


def extract_flattened_features(audio, sr, cfg=ADV_CONFIG, chroma="device"):
    """[R] _advanced.py:120-156 -> (290,) float64."""
    _m, f, status = process_batch_advanced(np.asarray(audio)[None], sr, cfg, chroma)
    _raise_failed(status)
    return f[0]


# ---------------------------------------------------------------------------
# on-disk layout (SURVEY.md 8a "layout"); normalisation stays in sklearn on the host
# ---------------------------------------------------------------------------
def save_processed_data1(out_dir, features, labels, metadata_df=None, config=BASIC_CONFIG):
    """[R] 1_preprocessing.py:295-343: impute, scale, np.save / pickle the same file names."""
    from sklearn.impute import SimpleImputer
    from sklearn.preprocessing import StandardScaler

    os.makedirs(out_dir, exist_ok=True)
    features = np.asarray(features, dtype=np.float64)
    features_clean = np.where(np.isinf(features), np.nan, features)
    imputer = SimpleImputer(strategy="mean", keep_empty_features=True)
    features_clean = imputer.fit_transform(features_clean)
    scaler = StandardScaler()
    features_normalized = scaler.fit_transform(features_clean)
    np.save(os.path.join(out_dir, "features_raw.npy"), features_clean)
    np.save(os.path.join(out_dir, "features_normalized.npy"), features_normalized)
    np.save(os.path.join(out_dir, "labels.npy"), np.asarray(labels))
    if metadata_df is not None:
        metadata_df.to_csv(os.path.join(out_dir, "metadata.csv"), index=False)
    for name, obj in (("scaler.pkl", scaler), ("imputer.pkl", imputer), ("config.pkl", dict(config))):
        with open(os.path.join(out_dir, name), "wb") as f:
            pickle.dump(obj, f)
    return features_clean, features_normalized


def save_processed_data2(out_dir, mel_spectrograms, flat_features, labels, lyrics_embeddings=None,
                         metadata_df=None, config=ADV_CONFIG, device_scaler=None):
    """[R] _advanced.py:376-421: scale the flattened mel images and the 290-vectors, save.

    ``device_scaler``: CUDA device ordinal to fit / apply the big (N, 131072) mel StandardScaler on
    the GPU; None keeps it in sklearn on the host.  Either way ``mel_scaler.pkl`` is
    a sklearn StandardScaler."""
    from sklearn.impute import SimpleImputer
    from sklearn.preprocessing import StandardScaler

    os.makedirs(out_dir, exist_ok=True)
    mel = np.asarray(mel_spectrograms, dtype=np.float32)
    n, h, w = mel.shape
    if device_scaler is None:
        mel_scaler = StandardScaler()
        mel_norm = mel_scaler.fit_transform(mel.reshape(n, -1)).reshape(n, h, w).astype(np.float32)
    else:
        import torch
        from .scaler import fit_transform_device

        xd = torch.from_numpy(np.ascontiguousarray(mel.reshape(n, -1))).to(torch.device("cuda", device_scaler))
        yd, mel_scaler = fit_transform_device(xd, inplace=True)
        mel_norm = yd.cpu().numpy().reshape(n, h, w)
    flat = np.asarray(flat_features, dtype=np.float64)
    flat = np.where(np.isinf(flat), np.nan, flat)
    imputer = SimpleImputer(strategy="mean", keep_empty_features=True)
    flat = imputer.fit_transform(flat)
    flat_scaler = StandardScaler()
    flat_norm = flat_scaler.fit_transform(flat)
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
