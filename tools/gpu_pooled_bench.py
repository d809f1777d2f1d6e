"""Times the pooled-columns-only path (fused dB + DCT + pooling) against the full outputs + pool kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import hybrid_language_music_clustering_vae_b200 as hl

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

B, n = 10000, 66150
ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max)
y = torch.randn((B, n), device="cuda") * 0.1
out = ex.extract_device(y, pooled=True)
full = timeit(lambda: ex.extract_device(y, out=out, pooled=True))
fused = timeit(lambda: ex.extract_pooled_device(y))
print(f"10000 x 3 s, 346 pooled columns: full outputs + pool kernel {full:.3f} ms ({B/full*1e3:.0f} clips/s); "
      f"fused pooled-only {fused:.3f} ms ({B/fused*1e3:.0f} clips/s)")
