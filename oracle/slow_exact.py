"""Independent float64 cross-check of the oracle (TEST INFRASTRUCTURE ONLY).

Nothing here shares code with ``librosa_oracle.py``: the DFT is an explicit
cos/sin matrix product, framing is an explicit loop, the mel filterbank is
written from the triangle definition, the DCT is an explicit cosine matrix.
It follows the published definitions in float64
throughout, so it agrees with the oracle to float32 rounding, not bit for bit.
Sizes must stay small (O(n_fft^2) per frame).
"""
from __future__ import annotations

import numpy as np


def _pad(y, pad, mode):
    n = len(y)
    out = np.zeros(n + 2 * pad)
    for i in range(n + 2 * pad):
        s = i - pad
        if 0 <= s < n:
            out[i] = y[s]
        elif mode == "constant":
            out[i] = 0.0
        elif mode == "edge":
            out[i] = y[min(max(s, 0), n - 1)]
        elif mode == "reflect":
            if n == 1:
                out[i] = y[0]
                continue
            period = 2 * (n - 1)
            r = s % period
            out[i] = y[r] if r < n else y[period - r]
        else:
            raise ValueError(mode)
    return out


def frames(y, n_fft, hop, center=True, mode="constant"):
    y = np.asarray(y, dtype=np.float64)
    yp = _pad(y, n_fft // 2, mode) if center else y
    T = 1 + (len(yp) - n_fft) // hop
    return np.stack([yp[t * hop: t * hop + n_fft] for t in range(T)], axis=1)  # (n_fft, T)


def hann(N):
    k = np.arange(N)
    return 0.5 - 0.5 * np.cos(2 * np.pi * k / N)


def stft(y, n_fft, hop, center=True, mode="constant"):
    fr = frames(y, n_fft, hop, center, mode) * hann(n_fft)[:, None]
    k = np.arange(n_fft // 2 + 1)[:, None]
    n = np.arange(n_fft)[None, :]
    ang = 2 * np.pi * k * n / n_fft
    return (np.cos(ang) - 1j * np.sin(ang)) @ fr


def mel_filterbank(sr, n_fft, n_mels):
    def hz2mel(f):
        return f / (200.0 / 3) if f < 1000.0 else 15.0 + np.log(f / 1000.0) / (np.log(6.4) / 27.0)

    def mel2hz(m):
        return m * (200.0 / 3) if m < 15.0 else 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0))

    pts = [mel2hz(hz2mel(0.0) + i * (hz2mel(sr / 2) - hz2mel(0.0)) / (n_mels + 1)) for i in range(n_mels + 2)]
    F = n_fft // 2 + 1
    fb = np.zeros((n_mels, F))
    for m in range(n_mels):
        lo, ce, hi = pts[m], pts[m + 1], pts[m + 2]
        for k in range(F):
            f = k * sr / n_fft
            if lo < f <= ce:
                w = (f - lo) / (ce - lo)
            elif ce < f < hi:
                w = (hi - f) / (hi - ce)
            else:
                w = 0.0
            fb[m, k] = w * 2.0 / (hi - lo)
    return fb


def power_to_db(S, ref, amin=1e-10, top_db=80.0):
    r = S.max() if ref == "max" else ref
    out = 10 * np.log10(np.maximum(amin, S)) - 10 * np.log10(max(amin, r))
    return np.maximum(out, out.max() - top_db)


def dct_ortho(x, n_out):
    N = x.shape[0]
    k = np.arange(n_out)[:, None]
    m = np.arange(N)[None, :]
    D = np.cos(np.pi * k * (2 * m + 1) / (2 * N)) * np.sqrt(2.0 / N)
    D[0] *= np.sqrt(0.5)
    return D @ x


def features(y, sr=22050, n_fft=512, hop=128, n_mels=40, n_mfcc=13, mode="constant", roll=0.85, thr=1e-10):
    X = stft(y, n_fft, hop, True, mode)
    S = np.abs(X)
    mel = mel_filterbank(sr, n_fft, n_mels) @ (S ** 2)
    out = {"S": S, "mel": mel, "logmel": power_to_db(mel, "max"),
           "mfcc": dct_ortho(power_to_db(mel, 1.0), n_mfcc)}
    freq = np.arange(n_fft // 2 + 1) * sr / n_fft
    tot = S.sum(axis=0)
    den = np.where(tot < np.finfo(np.float32).tiny, 1.0, tot)
    cen = (freq[:, None] * S).sum(axis=0) / den
    bw = np.sqrt((S * (freq[:, None] - cen[None, :]) ** 2).sum(axis=0) / den)
    cum = np.cumsum(S, axis=0)
    ro = np.array([freq[np.argmax(cum[:, t] >= roll * cum[-1, t])] for t in range(S.shape[1])])
    fe = frames(y, n_fft, hop, True, "edge")
    neg = fe < -thr
    zcr = (neg[1:] != neg[:-1]).sum(axis=0) / n_fft
    fr = frames(y, n_fft, hop, True, mode)
    rms = np.sqrt((fr ** 2).mean(axis=0))
    out["stats"] = np.stack([cen, bw, ro, zcr, rms])
    return out
