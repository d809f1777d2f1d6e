"""Worker body of bench.py's CPU arm (TEST / BENCH INFRASTRUCTURE ONLY, like the rest of ``oracle/``).

One task = a block of clips of a memory-mapped ``.npy`` file, each run through the oracle port of the
reference's ``extract_all_features`` ([R] src/1_preprocessing.py:105-129: five STFT-bearing librosa calls
+ zcr + rms; chroma left out, as in the GPU step it is compared with).  Lives in an importable module so
that joblib workers resolve it by name and keep the mapped file between tasks.
"""
from __future__ import annotations

import numpy as np

from . import librosa_oracle as orc

_MAPPED = {}


def block(path: str, lo: int, hi: int, sr: int = 22050) -> float:
    a = _MAPPED.get(path)
    if a is None:
        _MAPPED.clear()
        a = _MAPPED[path] = np.load(path, mmap_mode="r")
    acc = 0.0
    for i in range(lo, hi):
        f = orc.extract_all_features(np.asarray(a[i]), sr, with_chroma=False)
        acc += float(f[0])
    return acc
