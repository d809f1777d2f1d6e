import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import hybrid_language_music_clustering_vae_b200 as hl
from hybrid_language_music_clustering_vae_b200 import synth
B, n = 10000, 66150
ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max)
T = ex.num_frames(n)
h = torch.empty((B, n), dtype=torch.float32, pin_memory=True)
synth.synth_batch(B, n, seed=1, mixture=False, out=h.numpy())
out = {"logmel": torch.empty((B, 128, T), dtype=torch.float32, pin_memory=True).numpy(),
       "mfcc": torch.empty((B, 40, T), dtype=torch.float32, pin_memory=True).numpy(),
       "stats": torch.empty((B, 5, T), dtype=torch.float32, pin_memory=True).numpy(),
       "status": torch.empty((B,), dtype=torch.int32, pin_memory=True).numpy()}
def run(**kw):
    for _ in range(2): ex.extract_host(h.numpy(), out=out, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(8): ex.extract_host(h.numpy(), out=out, **kw)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 8
    print(kw, f"{dt*1e3:.2f} ms  {B/dt:.0f} clips/s  H2D {B*n*4/dt/1e9:.1f} GB/s", flush=True)
run()
for chunk in (64, 128, 256, 512, 1024, 2048):
    for ns in (2, 3, 4):
        run(chunk_clips=chunk, n_streams=ns)
