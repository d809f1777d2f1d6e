"""Shared parity helpers: run the oracle per clip and compare with the CUDA outputs
using the tolerances BASELINE.json's north_star states."""
from __future__ import annotations

import numpy as np

from oracle import librosa_oracle as orc

LOGMEL_TOL_DB = 0.01      # log-mel: max-abs <= 0.01 dB
REL_TOL = 1e-4            # MFCC and statistics: <= 1e-4 of the clip's max |reference|


def oracle_clip(y, *, sr=22050, n_fft=2048, hop_length=512, n_mels=128, n_mfcc=40, pad_mode="constant",
                center=True, win_length=None, window="hann", ref=np.max, top_db=80.0, power=2.0,
                roll_percent=0.85, **mel_kw):
    """`mel_kw`: htk / fmin / fmax / norm of librosa.filters.mel."""
    kw = dict(n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window, center=center,
              pad_mode=pad_mode)
    mel = orc.melspectrogram(y=y, sr=sr, n_mels=n_mels, power=power, **kw, **mel_kw)
    out = {"mel": mel, "logmel": orc.power_to_db(mel, ref=ref, top_db=top_db)}
    if n_mfcc:
        out["mfcc"] = orc.mfcc(y=y, sr=sr, n_mfcc=n_mfcc, n_mels=n_mels, power=power, **kw, **mel_kw)
    S = np.abs(orc.stft(y, **kw))
    out["S"] = S
    stats = np.empty((5, mel.shape[-1]), np.float64)
    stats[0] = orc.spectral_centroid(S=S, sr=sr, n_fft=n_fft)[0]
    stats[1] = orc.spectral_bandwidth(S=S, sr=sr, n_fft=n_fft)[0]
    stats[2] = orc.spectral_rolloff(S=S, sr=sr, n_fft=n_fft, roll_percent=roll_percent)[0]
    stats[3] = orc.zero_crossing_rate(y, frame_length=n_fft, hop_length=hop_length, center=center)[0]
    stats[4] = orc.rms(y=y, frame_length=n_fft, hop_length=hop_length, center=center, pad_mode=pad_mode)[0]
    out["stats"] = stats
    return out


def rolloff_margin_ok(S, sr, n_fft, got_hz, want_hz, roll_percent=0.85, margin=2e-6):
    """A differing rolloff bin is acceptable only where the oracle's own float32 cumulative
    sum sits within `margin` (relative) of the threshold at one of the two bins."""
    binhz = sr / n_fft
    cum = np.cumsum(S.astype(np.float64), axis=0)
    thr = roll_percent * cum[-1]
    ok = np.ones(S.shape[1], bool)
    for t in np.nonzero(np.abs(got_hz - want_hz) > 1e-3 * binhz)[0]:
        kg, kw = int(round(got_hz[t] / binhz)), int(round(want_hz[t] / binhz))
        lo, hi = min(kg, kw), max(kg, kw)
        # every bin in [lo, hi) must be a near tie
        rel = np.abs(cum[lo:hi, t] - thr[t]) / max(thr[t], 1e-30)
        ok[t] = bool(np.all(rel < margin))
    return ok


def compare_clip(got: dict, want: dict, *, sr=22050, n_fft=2048, roll_percent=0.85):
    """Returns a dict of error metrics; raises nothing."""
    m = {}
    m["frames_equal"] = got["logmel"].shape == want["logmel"].shape
    m["logmel_maxabs_db"] = float(np.abs(got["logmel"] - want["logmel"]).max())
    if "mfcc" in got and "mfcc" in want:
        scale = max(float(np.abs(want["mfcc"]).max()), 1e-6)
        m["mfcc_rel"] = float(np.abs(got["mfcc"] - want["mfcc"]).max() / scale)
    if "stats" in got:
        g, w = got["stats"].astype(np.float64), want["stats"]
        for i, name in enumerate(("centroid", "bandwidth", "rolloff", "zcr", "rms")):
            scale = max(float(np.abs(w[i]).max()), 1e-12)
            if name == "rolloff":
                bad = np.abs(g[i] - w[i]) > 1e-3 * sr / n_fft
                okm = rolloff_margin_ok(want["S"], sr, n_fft, g[i], w[i], roll_percent)
                m["rolloff_flips"] = int(bad.sum())
                m["rolloff_unexplained"] = int((bad & ~okm).sum())
            else:
                m[name + "_rel"] = float(np.abs(g[i] - w[i]).max() / scale)
    return m


def assert_clip(m: dict, where=""):
    assert m["frames_equal"], f"frame count differs {where}"
    assert m["logmel_maxabs_db"] <= LOGMEL_TOL_DB, f"log-mel {m['logmel_maxabs_db']} dB {where}"
    if "mfcc_rel" in m:
        assert m["mfcc_rel"] <= REL_TOL, f"mfcc {m['mfcc_rel']} {where}"
    for k in ("centroid_rel", "bandwidth_rel", "zcr_rel", "rms_rel"):
        if k in m:
            assert m[k] <= REL_TOL, f"{k} {m[k]} {where}"
    if "rolloff_unexplained" in m:
        assert m["rolloff_unexplained"] == 0, f"rolloff flips without a tie: {m} {where}"
