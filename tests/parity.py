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
        # librosa.feature.mfcc: DCT of power_to_db(melspectrogram) with ref=1.0, amin=1e-10, top_db=80
        out["mfcc"] = orc.mfcc(S=orc.power_to_db(mel), n_mfcc=n_mfcc)
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


FP32_FLOOR = 2.0 ** -22      # per-bin magnitude floor of a float32 FFT, relative to the frame's largest bin


def fp32_floor_allowance(S, sr, n_fft):
    """How far centroid / bandwidth can move when every bin of the oracle's magnitude spectrum is
    perturbed by FP32_FLOOR * max|X|.  The oracle runs librosa's float64 FFT; the device runs a
    float32 FFT (as north_star specifies), whose rounding floor sits ~130 dB under the strongest bin.
    For ordinary spectra the allowance is orders of magnitude below 1e-4; it only matters for
    single-line spectra (DC, on-bin tones, n=1 clips) where bandwidth = sqrt(sum S (f-c)^2 / sum S)
    weighs that floor by up to (sr/2)^2."""
    S = S.astype(np.float64)
    freq = np.arange(S.shape[0])[:, None] * (sr / n_fft)
    tot = np.maximum(S.sum(axis=0), 1e-300)
    a = FP32_FLOOR * S.max(axis=0)
    cen = (freq * S).sum(axis=0) / tot
    dev2 = (freq - cen[None, :]) ** 2
    bw2 = (S * dev2).sum(axis=0) / tot
    d_cen = a * np.abs(freq - cen[None, :]).sum(axis=0) / tot
    d_bw = np.sqrt(bw2 + a * dev2.sum(axis=0) / tot) - np.sqrt(bw2)
    return d_cen, d_bw


def is_degenerate(y) -> bool:
    """Clips whose spectrum is a single line or empty: constant (zero / DC), a single non-zero sample, or
    n <= 2.  Only for these may centroid / bandwidth use the float32-FFT floor allowance."""
    y = np.asarray(y)
    return bool(y.size <= 2 or np.ptp(y) == 0 or np.count_nonzero(y) <= 1)


def compare_clip(got: dict, want: dict, *, sr=22050, n_fft=2048, roll_percent=0.85, y=None):
    """Returns a dict of error metrics; raises nothing.  Pass the clip `y` so that the floor allowance is
    confined to degenerate clips; without it the raw errors are asserted."""
    m = {"degenerate": bool(y is not None and is_degenerate(y))}
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
            elif name in ("centroid", "bandwidth"):
                allow = fp32_floor_allowance(want["S"], sr, n_fft)[0 if name == "centroid" else 1]
                m[name + "_rel"] = float(np.maximum(np.abs(g[i] - w[i]) - allow, 0.0).max() / scale)
                m[name + "_raw_rel"] = float(np.abs(g[i] - w[i]).max() / scale)
            else:
                m[name + "_rel"] = float(np.abs(g[i] - w[i]).max() / scale)
    return m


def assert_clip(m: dict, where=""):
    assert m["frames_equal"], f"frame count differs {where}"
    assert m["logmel_maxabs_db"] <= LOGMEL_TOL_DB, f"log-mel {m['logmel_maxabs_db']} dB {where}"
    if "mfcc_rel" in m:
        assert m["mfcc_rel"] <= REL_TOL, f"mfcc {m['mfcc_rel']} {where}"
    for k in ("centroid", "bandwidth"):
        # raw error for every ordinary clip; the float32-floor allowance only for single-line spectra
        key = k + ("_rel" if m.get("degenerate") else "_raw_rel")
        if key in m:
            assert m[key] <= REL_TOL, f"{key} {m[key]} {where}"
    for k in ("zcr_rel", "rms_rel"):
        if k in m:
            assert m[k] <= REL_TOL, f"{k} {m[k]} {where}"
    if "rolloff_unexplained" in m:
        assert m["rolloff_unexplained"] == 0, f"rolloff flips without a tie: {m} {where}"
