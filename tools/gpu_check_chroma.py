"""Diagnostic: chroma_stft / estimate_tuning on device vs the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import hybrid_language_music_clustering_vae_b200 as hl
from oracle import librosa_oracle as orc

for n in (22050, 66150):
    y = hl.synth.synth_batch(24, n, seed=31)
    kinds = hl.synth.mixture_kinds(24)
    ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max)
    out = ex.extract_device(torch.from_numpy(y).cuda(), chroma=True, pooled=True)
    torch.cuda.synchronize()
    ch = out["chroma"].cpu().numpy(); tu = out["tuning"].cpu().numpy()
    print("pooled", out["pooled"].shape)
    for i in range(len(y)):
        S = np.abs(orc.stft(y[i])) ** 2
        p, m = orc.piptrack(S=S, sr=22050, n_fft=2048)
        t_or = orc.estimate_tuning(S=S, sr=22050, n_fft=2048, bins_per_octave=12)
        c_or = orc.chroma_stft(y=y[i], sr=22050)
        c_same = orc.chroma_stft(y=y[i], sr=22050, tuning=float(tu[i]))
        print(f"{i:2d} {kinds[i]:10s} cand={int((p>0).sum()):6d} tuning dev={tu[i]:+.2f} oracle={t_or:+.2f}  "
              f"chroma diff(oracle tuning)={np.abs(ch[i]-c_or).max():.2e}  diff(same tuning)={np.abs(ch[i]-c_same).max():.2e}")
