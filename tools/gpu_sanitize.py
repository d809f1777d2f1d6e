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
# round 2: the front end (odd lengths, every channel / format variant, ragged clips), the fixed image and the normalisers
rng = np.random.default_rng(0)
for sr_in, ch, dt in ((44100, 2, np.int16), (48000, 1, np.int16), (22050, 3, np.int16), (16000, 2, np.float32),
                      (22050, 1, np.float32), (11025, 1, np.int16)):
    n = 2999 + ch
    raw = rng.integers(-9000, 9000, size=(3, n, ch)).astype(np.int16)
    raw = raw if dt == np.int16 else (raw / 32768.0).astype(np.float32)
    valid = np.array([n, n // 2 + 1, 1], np.int64)
    w = ex.load_frontend_device(torch.from_numpy(raw).cuda(), sr_in=sr_in, valid_frames=valid)
    r = ex.extract_host(raw, sr_in=sr_in, valid_frames=valid, pad_to=int(np.ceil(n * 22050 / sr_in)) + 77,
                        pooled=True, chroma="pooled", fixed_frames=40, wave_out=True, chunk_clips=2, n_streams=2)
    torch.cuda.synchronize()
    assert torch.isfinite(w).all() and np.isfinite(r["pooled"]).all() and np.isfinite(r["fixed_logmel"]).all(), (sr_in, ch)
from hybrid_language_music_clustering_vae_b200.scaler import fit_transform_device, fit_transform_tabular_device
for N, D in ((1, 5), (300, 370), (257, 33)):
    x = torch.randn((N, D), device="cuda", dtype=torch.float64)
    if N > 1:
        x[0, 0] = float("nan"); x[1, 1] = float("inf"); x[:, 2] = float("nan")
    imp, sc, _a, _b = fit_transform_tabular_device(x)
    assert torch.isfinite(sc).all()
    fit_transform_device(torch.randn((N, 4 * D), device="cuda"))
db = hl.power_to_db(torch.rand((70000, 2, 3), device="cuda") + 0.1, ref=np.max)
torch.cuda.synchronize()
print("sanitize run ok:", len(cases) + 1, "plans + front end + normalisers")
