import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import hybrid_language_music_clustering_vae_b200 as hl
from parity import oracle_clip, compare_clip
def run(B, n, **kw):
    y = hl.synth.synth_batch(B, n, seed=11)
    ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max, **kw)
    out = ex.extract_device(torch.from_numpy(y).cuda()); torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in out.items()}
    worst = {}
    for i in range(B):
        want = oracle_clip(y[i], n_mfcc=40, **kw)
        got = {k: out[k][i] for k in ("logmel", "mfcc", "stats")}
        m = compare_clip(got, want, n_fft=kw.get("n_fft", 2048), sr=22050)
        for k, v in m.items():
            if isinstance(v, bool): worst[k] = worst.get(k, True) and v
            else: worst[k] = max(worst.get(k, 0), v)
    print(f"B={B} n={n} fast={ex.uses_fast_path()} {kw}:", {k: (f"{v:.3g}" if isinstance(v, float) else v) for k, v in worst.items()}, flush=True)
run(8, 30000, n_fft=4096, hop_length=1024)
run(8, 66150, n_fft=4096, hop_length=1024, pad_mode="reflect")
run(4, 4095, n_fft=4096, hop_length=1024)
run(4, 66150, n_fft=4096, hop_length=777, center=False)
run(6, 66150, n_fft=4096, hop_length=1024, power=1.0)
ex = hl.FeatureExtractor(n_fft=4096, hop_length=1024, n_mfcc=40, ref=np.max)
y = torch.randn(4096, 66150, device="cuda") * 0.1
out = ex.extract_device(y)
for _ in range(2): ex.extract_device(y, out=out)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(3): ex.extract_device(y, out=out)
torch.cuda.synchronize(); dt = (time.time() - t0) / 3
print(f"4096 clips n_fft=4096: {dt*1e3:.2f} ms -> {4096/dt:.0f} clips/s")
