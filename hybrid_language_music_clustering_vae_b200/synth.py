"""Synthetic 22.05 kHz mono clips of the shapes BASELINE.json names.

Mixture per batch: 40 % white Gaussian (sigma 0.1), 30 % "music-like" decaying
harmonic stacks, 15 % full-scale uniform, 10 % half-silent (second half exact
zeros, mirroring the reference's zero padding, [R] src/1_preprocessing.py:148),
5 % edge clips (all-zero / DC / single impulse).
"""
from __future__ import annotations

import numpy as np

SR = 22050


def synth_clip(kind: str, n: int, rng: np.random.Generator, sr: int = SR) -> np.ndarray:
    t = np.arange(n, dtype=np.float64) / sr
    if kind == "white":
        y = 0.1 * rng.standard_normal(n)
    elif kind == "harmonic":
        f0 = rng.uniform(80.0, 880.0)
        y = np.zeros(n)
        k = 1
        while k * f0 < sr / 2 and k <= 40:
            y += np.sin(2 * np.pi * k * f0 * t + rng.uniform(0, 2 * np.pi)) / k
            k += 1
        y *= np.exp(-t * rng.uniform(0.2, 3.0))
        y = 0.5 * y / max(1e-9, np.abs(y).max()) + 1e-4 * rng.standard_normal(n)
    elif kind == "uniform":
        y = rng.uniform(-1.0, 1.0, n)
    elif kind == "halfsilent":
        y = 0.1 * rng.standard_normal(n)
        y[n // 2:] = 0.0
    elif kind == "zero":
        y = np.zeros(n)
    elif kind == "dc":
        y = np.full(n, 0.25)
    elif kind == "impulse":
        y = np.zeros(n)
        y[int(rng.integers(0, n))] = 1.0
    else:
        raise ValueError(kind)
    return np.clip(y, -1.0, 1.0).astype(np.float32)


def mixture_kinds(B: int):
    kinds = []
    for i in range(B):
        u = (i * 0.6180339887498949) % 1.0   # low-discrepancy, deterministic
        if u < 0.40:
            kinds.append("white")
        elif u < 0.70:
            kinds.append("harmonic")
        elif u < 0.85:
            kinds.append("uniform")
        elif u < 0.95:
            kinds.append("halfsilent")
        else:
            kinds.append(("zero", "dc", "impulse")[i % 3])
    return kinds


def synth_batch(B: int, n: int, seed: int = 20260, mixture: bool = True, out=None) -> np.ndarray:
    rng = np.random.default_rng(seed)
    y = out if out is not None else np.empty((B, n), dtype=np.float32)
    if not mixture:
        # throughput is data independent: white noise only, generated in blocks
        blk = max(1, (64 << 20) // (4 * n))
        for i in range(0, B, blk):
            j = min(B, i + blk)
            y[i:j] = (0.1 * rng.standard_normal((j - i, n), dtype=np.float32))
        return y
    for i, kind in enumerate(mixture_kinds(B)):
        y[i] = synth_clip(kind, n, rng)
    return y
