"""Measures the BASELINE.json configs that are not the bench line (parity cases) for the record:
  configs[2]  30-s clips (GTZAN shape), one GPU's share (125 clips) and 1000 clips
  configs[4]  batch 1..65536 x n_fft 512/1024/2048/4096, hop = n_fft/4, 3-s clips: latency and throughput
Writes one JSON document to stdout."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import hybrid_language_music_clustering_vae_b200 as hl

def timeit(fn, reps, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

res = {"gpu": torch.cuda.get_device_name(0), "configs2_30s": [], "configs4_sweep": []}
for B in (125, 1000):
    n = 661500
    ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max)
    y = torch.randn((B, n), device="cuda") * 0.1
    out = ex.extract_device(y)
    ms = timeit(lambda: ex.extract_device(y, out=out), 5)
    res["configs2_30s"].append({"clips": B, "ms": ms, "clips_per_s": B / ms * 1e3, "audio_hours_per_s": B * 30 / 3600 / ms * 1e3})
    del y, out
for n_fft in (512, 1024, 2048, 4096):
    ex = hl.FeatureExtractor(n_fft=n_fft, hop_length=n_fft // 4, n_mfcc=40, ref=np.max)
    for B in (1, 8, 64, 512, 4096, 65536 if n_fft == 2048 else 16384):
        y = torch.randn((B, 66150), device="cuda") * 0.1
        out = ex.extract_device(y)
        ms = timeit(lambda: ex.extract_device(y, out=out), 20 if B <= 512 else 3)
        row = {"n_fft": n_fft, "batch": B, "register_fft_kernel": ex.uses_fast_path(),
               "latency_us": ms * 1e3, "clips_per_s": B / ms * 1e3}
        if B <= 64:          # launch-bound: the same call replayed as a CUDA graph
            g = ex.capture_device(y)
            row["graph_latency_us"] = timeit(g.replay, 50) * 1e3
            g.close()
        res["configs4_sweep"].append(row)
        del y, out
print(json.dumps(res, indent=1))
