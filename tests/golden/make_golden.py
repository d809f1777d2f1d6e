"""Regenerates tests/golden/golden_v1.npz.

The reference's arithmetic lives in librosa, which cannot be imported in this image, and the
reference holds no golden vectors of its own.  These vectors are therefore
outputs of the CPU oracle (oracle/librosa_oracle.py) on seeded inputs; tests/test_golden.py
checks them against the oracle (drift guard), against the independent direct-DFT
implementation (oracle/slow_exact.py), and -- on the GPU box -- against the CUDA path.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from hybrid_language_music_clustering_vae_b200_synth_shim import synth_clip  # noqa: E402
from parity import oracle_clip  # noqa: E402

CASES = {
    "a": dict(n=6000, kinds=["harmonic", "white", "halfsilent", "dc", "impulse", "zero"],
              kw=dict(n_fft=2048, hop_length=512, n_mels=128, n_mfcc=40, pad_mode="constant")),
    "b": dict(n=3000, kinds=["harmonic", "uniform"],
              kw=dict(n_fft=512, hop_length=128, n_mels=40, n_mfcc=13, pad_mode="reflect")),
}


def main():
    out = {}
    for name, c in CASES.items():
        rng = np.random.default_rng(1234 + ord(name))
        y = np.stack([synth_clip(k, c["n"], rng) for k in c["kinds"]])
        out[f"{name}_y"] = y
        res = [oracle_clip(y[i], **c["kw"]) for i in range(len(y))]
        for key in ("logmel", "mfcc", "stats"):
            out[f"{name}_{key}"] = np.stack([r[key] for r in res]).astype(np.float32 if key != "stats" else np.float64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
