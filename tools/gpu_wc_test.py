"""Does write-combined pinned host memory speed up the H2D leg?  (cudaHostAllocWriteCombined vs default pinned)"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hybrid_language_music_clustering_vae_b200 as hl
from hybrid_language_music_clustering_vae_b200 import synth

rt = C.CDLL("libcudart.so.12")
B, n = 10000, 66150
nbytes = B * n * 4

def host_alloc(flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags)) == 0
    return p, np.ctypeslib.as_array((C.c_float * (B * n)).from_address(p.value)).reshape(B, n)

torch.zeros(1).cuda()
d = torch.empty((B, n), dtype=torch.float32, device="cuda")
ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max)
T = ex.num_frames(n)
out = {"logmel": torch.empty((B, 128, T), dtype=torch.float32, pin_memory=True).numpy(),
       "mfcc": torch.empty((B, 40, T), dtype=torch.float32, pin_memory=True).numpy(),
       "stats": torch.empty((B, 5, T), dtype=torch.float32, pin_memory=True).numpy(),
       "status": torch.empty((B,), dtype=torch.int32, pin_memory=True).numpy()}
for name, flags in (("default pinned", 0), ("write-combined", 4)):
    p, arr = host_alloc(flags)
    src = synth.synth_batch(512, n, seed=1, mixture=False)
    for lo in range(0, B, 512):
        arr[lo:lo + 512] = src[: min(512, B - lo)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(2):
        e0.record()
        for _ in range(3):
            assert rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr()), p, C.c_size_t(nbytes), C.c_int(1), C.c_void_p(st)) == 0
        e1.record(); torch.cuda.synchronize()
    bare = 3 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    for _ in range(2): ex.extract_host(arr, out=out)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(6): ex.extract_host(arr, out=out)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 6
    print(f"{name}: bare H2D {bare:.1f} GB/s; e2e {dt*1e3:.2f} ms = {B/dt:.0f} clips/s (H2D {nbytes/dt/1e9:.1f} GB/s)", flush=True)
    rt.cudaFreeHost(p)
