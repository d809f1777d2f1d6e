"""Regenerates tests/golden/torchaudio_v1.npz: outputs of torchaudio's librosa-compatible transforms on seeded
clips.  Unlike golden_v1.npz these vectors do NOT come from our own oracle: torchaudio is an independent code
base whose MelSpectrogram(norm="slaney", mel_scale="slaney") / MFCC(log_mels=False) / spectral_centroid are
written (and tested upstream) against librosa.  The fixtures let the GPU suite check the CUDA path against a
third party without importing anything at run time.

    python tests/golden/make_torchaudio_golden.py            (needs torch + torchaudio, CPU is enough)
"""
import os
import sys

import numpy as np
import torch
import torchaudio as ta

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hybrid_language_music_clustering_vae_b200_synth_shim import synth_clip  # noqa: E402

SR, N = 22050, 22050
KINDS = ["harmonic", "harmonic", "white", "uniform", "halfsilent"]


def main():
    rng = np.random.default_rng(777)
    y = np.stack([synth_clip(k, N, rng) for k in KINDS]).astype(np.float32)
    yt = torch.from_numpy(y).double()          # float64 transforms: the fixture's own rounding stays below 1e-7
    mk = dict(n_fft=2048, hop_length=512, n_mels=128, power=2.0, center=True, pad_mode="constant", norm="slaney",
              mel_scale="slaney")
    mel = ta.transforms.MelSpectrogram(sample_rate=SR, **mk).double()(yt)
    # clip by clip: amplitude_to_DB takes its top_db maximum over every leading dimension, librosa.feature.mfcc
    # (called per file by the scripts) over one clip
    tr = ta.transforms.MFCC(sample_rate=SR, n_mfcc=40, dct_type=2, norm="ortho", log_mels=False, melkwargs=mk).double()
    mfcc = torch.stack([tr(yt[i]) for i in range(len(y))])
    cen = ta.functional.spectral_centroid(yt, SR, pad=0, window=torch.hann_window(2048, dtype=torch.float64), n_fft=2048,
                                          hop_length=512, win_length=2048)      # reflect-padded spectrogram
    out = dict(y=y, mel=mel.numpy(), mfcc=mfcc.numpy(), centroid_reflect=cen.numpy(),
               versions=np.array([torch.__version__, ta.__version__]))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "torchaudio_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
