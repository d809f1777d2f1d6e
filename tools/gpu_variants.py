"""Device-resident timing of the n_fft = 2048 kernel variants (default, chroma, Hamming window, 40 / 64 / 256 mels) and
of the other register-FFT sizes: one line each, 10,000 x 3 s clips (CUDA events around extract_device)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hybrid_language_music_clustering_vae_b200 as hl

def timeit(fn, reps=8, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

B, n = 10000, 66150
y = torch.randn((B, n + 2), device="cuda")[:, :n] * 0.1
tag = " ".join(f"{k}={os.environ[k]}" for k in ("HLMC_NO_TMEM", "HLMC_NO_PREF") if k in os.environ) or "default"
cases = [("default", dict(), dict()), ("chroma", dict(), dict(chroma=True)), ("hamming", dict(window="hamming"), dict()),
         ("n_mels=40", dict(n_mels=40), dict()), ("n_mels=64", dict(n_mels=64), dict()), ("n_mels=256", dict(n_mels=256), dict()),
         ("n_fft=1024", dict(n_fft=1024, hop_length=256), dict()), ("n_fft=512", dict(n_fft=512, hop_length=128), dict()),
         ("n_fft=4096", dict(n_fft=4096, hop_length=1024), dict())]
for name, kw, ckw in cases:
    ex = hl.FeatureExtractor(n_mfcc=min(40, kw.get("n_mels", 128)), ref=np.max, **kw)
    out = ex.extract_device(y, **ckw)
    ms = timeit(lambda: ex.extract_device(y, out=out, **ckw))
    print(f"[{tag}] {name:12s} {ms:8.3f} ms per 10,000 clips  {B / ms * 1e3 / 1e6:6.3f} M clips/s", flush=True)
    del ex, out
