import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import hybrid_language_music_clustering_vae_b200 as hl
def timeit(fn, reps, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max)
for path in (0,):
    ex.set_path(path)
    for B in (1, 8, 64, 128, 256, 512):
        y = torch.randn((B, 66150), device="cuda") * 0.1
        out = ex.extract_device(y)
        ms = timeit(lambda: ex.extract_device(y, out=out), 50)
        g = ex.capture_device(y); gms = timeit(g.replay, 50); g.close()
        print(f"path {path} batch {B}: {ms*1e3:.1f} us, graph {gms*1e3:.1f} us")
