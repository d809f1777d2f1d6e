"""Times the frames kernel and the dB/DCT kernel of one 10,000 x 3 s device-resident step (events recorded on the
launching stream inside the C ABI).  The kernel variant is chosen by environment switches read at first launch
(HLMC_NO_MMA=1, HLMC_NO_PREF=1), so run one process per variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hybrid_language_music_clustering_vae_b200 as hl

B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 66150
ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max)
y = torch.randn((B, n + 2), device="cuda")[:, :n] * 0.1
out = ex.extract_device(y)
for _ in range(3): ex.extract_device(y, out=out)
torch.cuda.synchronize()
ex.set_timing(True)
for _ in range(10): ex.extract_device(y, out=out)
torch.cuda.synchronize()
f, d, k = ex.read_timing()
ex.set_timing(False)
tag = " ".join(f"{k_}={os.environ[k_]}" for k_ in ("HLMC_NO_MMA", "HLMC_NO_PREF") if k_ in os.environ) or "default"
print(f"[{tag}] B={B} n={n}: frames {f / k:.4f} ms  db_dct {d / k:.4f} ms  step {(f + d) / k:.4f} ms  -> {B / ((f + d) / k) * 1e3 / 1e6:.3f} M clips/s")
