"""Tiny batches through every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool memcheck python tools/gpu_sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hybrid_language_music_clustering_vae_b200 as hl

y = hl.synth.synth_batch(3, 9000, seed=5)
yd = torch.from_numpy(y).cuda()
cases = [dict(), dict(window="hamming"), dict(pad_mode="reflect", center=True), dict(center=False),
         dict(n_fft=1024, hop_length=256), dict(n_fft=512, hop_length=128), dict(n_fft=4096, hop_length=1024),
         dict(n_fft=256, hop_length=64)]
for kw in cases:
    ex = hl.FeatureExtractor(n_mfcc=20, ref=np.max, **kw)
    out = ex.extract_device(yd, pooled=True)
    torch.cuda.synchronize()
    assert torch.isfinite(out["logmel"]).all() and torch.isfinite(out["pooled"]).all(), kw
    ex.close()
ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max)
out = ex.extract_device(yd, chroma=True, pooled=True)
po = ex.extract_pooled_device(yd, chroma=True)
h = ex.extract_host(y, pooled=True, chunk_clips=2, n_streams=2)
torch.cuda.synchronize()
assert np.isfinite(h["pooled"]).all() and torch.isfinite(po["pooled"]).all()
print("sanitize run ok:", len(cases) + 1, "plans")
