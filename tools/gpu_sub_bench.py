import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import hybrid_language_music_clustering_vae_b200 as hl
def timeit(fn, reps=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n_fft in (512, 1024, 4096):
    ex = hl.FeatureExtractor(n_fft=n_fft, hop_length=n_fft // 4, n_mfcc=40, ref=np.max)
    y = torch.randn((8192, 66150), device="cuda") * 0.1
    out = ex.extract_device(y)
    ms = timeit(lambda: ex.extract_device(y, out=out))
    print(n_fft, f"{ms:.2f} ms  {8192/ms*1e3:.0f} clips/s")
