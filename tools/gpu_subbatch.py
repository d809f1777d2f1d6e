"""Does a mel-power scratch that stays in L2 pay?  The same 10,000 x 3 s step run as sub-batches of C clips (the
stream-ordered pool hands every call the same scratch block, so for small C the frames kernel's writes are still in
L2 when db_dct reads them and are overwritten before they are evicted)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hybrid_language_music_clustering_vae_b200 as hl

B, n = 10000, 66150
ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max)
y = torch.randn((B, n + 2), device="cuda")[:, :n] * 0.1
out = ex.extract_device(y)
def run(C):
    for s in range(0, B, C):
        sub = {k: v[s:s + C] for k, v in out.items()}
        ex.extract_device(y[s:s + C], out=sub)
for C in (10000, 5000, 2500, 1000, 500, 250):
    for _ in range(2): run(C)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run(C)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ex.set_timing(True); run(C); torch.cuda.synchronize(); f, d, k = ex.read_timing(); ex.set_timing(False)
    print(f"sub-batches of {C:5d}: step {ms:.3f} ms  (frames {f:.3f} + db_dct {d:.3f} ms over {k} calls)  {B / ms * 1e3 / 1e6:.3f} M clips/s", flush=True)
