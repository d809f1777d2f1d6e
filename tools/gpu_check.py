"""Diagnostic (not a test): prints parity metrics of the CUDA path against the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import hybrid_language_music_clustering_vae_b200 as hl
from parity import oracle_clip, compare_clip

def run(B, n, generic=False, **kw):
    y = hl.synth.synth_batch(B, n, seed=11)
    okw = dict(kw)
    ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max, **kw)
    if generic: ex.force_generic(True)
    out = ex.extract_device(torch.from_numpy(y).cuda())
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in out.items()}
    worst = {}
    for i in range(B):
        want = oracle_clip(y[i], n_mfcc=40, **okw)
        got = {k: out[k][i] for k in ("logmel", "mfcc", "stats")}
        m = compare_clip(got, want, n_fft=kw.get("n_fft", 2048), sr=22050)
        for k, v in m.items():
            if isinstance(v, bool): worst[k] = worst.get(k, True) and v
            else: worst[k] = max(worst.get(k, 0), v)
    print(f"B={B} n={n} generic={generic} fast={ex.uses_fast_path()} {kw}:")
    print("   ", {k: (f"{v:.3g}" if isinstance(v, float) else v) for k, v in worst.items()}, "status", out["status"].tolist()[:8], flush=True)

print(torch.cuda.get_device_name(0))
for gen in (True, False):
    run(20, 22050, generic=gen)
    run(6, 66150, generic=gen, pad_mode="reflect")
    run(4, 2047, generic=gen)
    run(3, 511, generic=gen, pad_mode="reflect")
run(4, 30000, n_fft=1024, hop_length=256)
run(4, 30000, n_fft=512, hop_length=128)
run(4, 30000, n_fft=4096, hop_length=1024)
run(6, 66150, n_fft=4096, hop_length=1024, pad_mode="reflect")
run(4, 30000, n_fft=4096, hop_length=1024, generic=True)
run(2, 661500)
run(40, 66150)                      # the mixture DESIGN.md quotes
run(8, 66150, window="hamming")     # window table path (no synthesised Hann, no early copy)
run(8, 66150, n_mels=40)
run(8, 66150, power=1.0)
# round 2: librosa.load's arithmetic on the device against the oracle (scipy.signal.resample_poly) ...
from oracle import librosa_oracle as orc
ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max)
rng = np.random.default_rng(5)
for sr_in, ch in ((44100, 2), (48000, 1), (16000, 1), (22050, 2), (8000, 3)):
    n = int(1.5 * sr_in) + 7
    t = np.arange(n) / sr_in
    x = np.stack([0.3 * np.sin(2 * np.pi * 330.0 * (c + 1) * t) + 0.05 * rng.standard_normal(n) for c in range(ch)], axis=1)
    pcm = np.clip(np.rint(x * 32768.0), -32768, 32767).astype(np.int16)[None]
    got = ex.load_frontend_device(torch.from_numpy(pcm).cuda(), sr_in=sr_in).cpu().numpy()[0]
    want = orc.load_pcm16(pcm[0], sr_in, sr=22050)[0]
    print(f"front end {sr_in} Hz x{ch} int16 -> 22050 Hz mono: {len(got)} samples, max abs diff vs scipy resample_poly {np.abs(got - want).max():.3g}", flush=True)
# ... and the tabular normalisation against sklearn
from sklearn.impute import SimpleImputer
from sklearn.preprocessing import StandardScaler
from hybrid_language_music_clustering_vae_b200.scaler import fit_transform_tabular_device, fit_transform_device
X = rng.standard_normal((1336, 370)) * rng.uniform(0.1, 40.0, 370) + rng.uniform(-100, 100, 370)
X[rng.integers(0, 1336, 40), rng.integers(0, 370, 40)] = np.nan
X[rng.integers(0, 1336, 10), rng.integers(0, 370, 10)] = np.inf
imp, scaled, _im, _sc = fit_transform_tabular_device(torch.from_numpy(X).cuda())
wi = SimpleImputer(strategy="mean").fit_transform(np.where(np.isinf(X), np.nan, X))
ws = StandardScaler().fit_transform(wi)
print(f"tabular (1336, 370) f64: imputed max abs diff {np.abs(imp.cpu().numpy() - wi).max():.3g}, scaled max abs diff {np.abs(scaled.cpu().numpy() - ws).max():.3g}")
M = (rng.standard_normal((200, 131072)) * 12 - 40).astype(np.float32)
ym, _ = fit_transform_device(torch.from_numpy(M).cuda())
print(f"mel scaler (200, 131072) f32: max abs diff vs sklearn {np.abs(ym.cpu().numpy() - StandardScaler().fit_transform(M)).max():.3g}", flush=True)
# throughput quick look
ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max)
y = torch.randn(2000, 66152, device="cuda")[:, :66150] * 0.1
for _ in range(2):
    out = ex.extract_device(y)
torch.cuda.synchronize()
t0 = time.time()
for _ in range(3):
    out = ex.extract_device(y, out=out)
torch.cuda.synchronize()
dt = (time.time() - t0) / 3
print(f"2000 clips x 3 s: {dt*1e3:.2f} ms -> {2000/dt:.0f} clips/s")
print("fp32 peak TFLOP/s", hl.measure_fp32_peak(0))
