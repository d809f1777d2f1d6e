"""Single-process multi-GPU extraction with the host gather (SURVEY 8e / north_star: "per-rank streams and a host
gather, no NCCL"): one pinned host array in, one host array out, one thread + plan per device.

  python tools/gpu_multi_gather.py [clips_total] > profiles/<tag>.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hybrid_language_music_clustering_vae_b200 as hl  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
n = 66150
ndev = torch.cuda.device_count()
h = torch.empty((B, n), dtype=torch.float32, pin_memory=True)
hl.synth.synth_batch(B, n, seed=20264, mixture=False, out=h.numpy())
pcm = torch.empty((B, n), dtype=torch.int16, pin_memory=True)
for lo in range(0, B, 1024):
    pcm.numpy()[lo:lo + 1024] = np.clip(np.rint(h.numpy()[lo:lo + 1024] * 32768.0), -32768, 32767).astype(np.int16)
kw = dict(n_mfcc=40, ref=np.max)
res = {"clips": B, "devices": ndev, "what": "sharding.extract_multi_gpu: one process, one thread + plan per device, "
       "outputs written into slices of shared host arrays"}
for name, x, opts in (
        ("basic_contract_f32", h.numpy(), dict(logmel=False, mfcc=False, stats=False, pooled=True, chroma="pooled")),
        ("basic_contract_pcm16", pcm.numpy(), dict(logmel=False, mfcc=False, stats=False, pooled=True, chroma="pooled")),
        ("full_f32", h.numpy(), dict())):
    for devs in sorted({1, min(2, ndev), min(4, ndev), ndev}):
        devices = list(range(devs))
        hl.sharding.extract_multi_gpu(x, devices, kw, **opts)          # plans, slots, pinned pages warm (a full pass)
        t0 = time.perf_counter()
        out = hl.sharding.extract_multi_gpu(x, devices, kw, **opts)
        dt = time.perf_counter() - t0
        res[f"{name}_x{devs}"] = {"clips_per_s": B / dt, "seconds": dt, "status_bad": int(out["status"].sum())}
        del out
print(json.dumps(res))
