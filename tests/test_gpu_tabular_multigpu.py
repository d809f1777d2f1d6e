"""SURVEY 8f-4 (inf -> nan, SimpleImputer, StandardScaler on device) and the multi-GPU paths of 8(e):
the cross-GPU statistics under a real 2-rank NCCL group and the single-process multi-device host gather.
The >1-GPU tests skip on a one-GPU box."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _messy(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)) * rng.uniform(0.1, 40.0, d) + rng.uniform(-100, 100, d)
    x[rng.integers(0, n, 30), rng.integers(0, d, 30)] = np.nan
    x[rng.integers(0, n, 10), rng.integers(0, d, 10)] = np.inf
    x[rng.integers(0, n, 10), rng.integers(0, d, 10)] = -np.inf
    x[:, 5] = 3.25                        # constant column: scale_ must become 1
    return x


def _sklearn_chain(x):
    from sklearn.impute import SimpleImputer
    from sklearn.preprocessing import StandardScaler

    clean = np.where(np.isinf(x), np.nan, x)
    im = SimpleImputer(strategy="mean")
    imputed = im.fit_transform(clean)
    sc = StandardScaler()
    return imputed, sc.fit_transform(imputed), im, sc


@pytest.mark.parametrize("n,d", [(1336, 370), (1336, 290), (1, 370), (70000, 33)])
def test_tabular_normalisation_matches_sklearn(built, n, d):
    import torch
    from hybrid_language_music_clustering_vae_b200.scaler import fit_transform_tabular_device

    x = _messy(n, d, n + d)
    if n > 10:
        x[:, 9] = np.nan                  # a column without any observed value: sklearn drops it
    want_imp, want_scaled, im, sc = _sklearn_chain(x)
    imp, scaled, im2, sc2 = fit_transform_tabular_device(torch.from_numpy(x).cuda())
    imp, scaled = imp.cpu().numpy(), scaled.cpu().numpy()
    assert imp.shape == want_imp.shape and scaled.shape == want_scaled.shape
    tol = 1e-12 * max(1.0, np.abs(want_imp).max())
    assert np.abs(imp - want_imp).max() <= tol
    assert np.allclose(scaled, want_scaled, rtol=1e-10, atol=1e-10)
    assert np.allclose(im2.statistics_, im.statistics_, rtol=1e-13, equal_nan=True)
    assert np.allclose(sc2.mean_, sc.mean_, rtol=1e-12, atol=tol) and np.allclose(sc2.var_, sc.var_, rtol=1e-10)
    assert np.array_equal(sc2.scale_ == 1.0, sc.scale_ == 1.0)
    # the rebuilt sklearn objects behave like fitted ones (they are what gets pickled)
    clean = np.where(np.isinf(x), np.nan, x)
    assert np.allclose(sc2.transform(im2.transform(clean)), want_scaled, rtol=1e-10, atol=1e-10)


_NCCL_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
from hybrid_language_music_clustering_vae_b200.scaler import fit_transform_device, fit_transform_tabular_device
from hybrid_language_music_clustering_vae_b200.sharding import shard_bounds
from sklearn.preprocessing import StandardScaler
from sklearn.impute import SimpleImputer
rng = np.random.default_rng(0)
N, D = 301, 4096
X = (rng.standard_normal((N, D)) * 7 - 30).astype(np.float32)          # same matrix on every rank
lo, hi = shard_bounds(N, rank, world)
y, sc = fit_transform_device(torch.from_numpy(X[lo:hi]).to(dev))
ref = StandardScaler().fit(X)
assert sc.n_samples_seen_ == N
assert np.allclose(sc.mean_, ref.mean_, rtol=1e-6, atol=1e-5) and np.allclose(sc.var_, ref.var_, rtol=1e-5)
assert np.allclose(y.cpu().numpy(), ref.transform(X)[lo:hi], rtol=1e-4, atol=1e-4)
T = rng.standard_normal((N, 370)) * 9 + 4
T[rng.integers(0, N, 40), rng.integers(0, 370, 40)] = np.nan
T[7, 3] = np.inf
imp, scaled, im, sc2 = fit_transform_tabular_device(torch.from_numpy(T[lo:hi]).to(dev))
clean = np.where(np.isinf(T), np.nan, T)
want_imp = SimpleImputer(strategy="mean").fit_transform(clean)
want = StandardScaler().fit_transform(want_imp)
assert np.allclose(imp.cpu().numpy(), want_imp[lo:hi], rtol=1e-12, atol=1e-12)
assert np.allclose(scaled.cpu().numpy(), want[lo:hi], rtol=1e-10, atol=1e-10)
dist.barrier()
if rank == 0:
    print("NCCL_SCALER_OK")
dist.destroy_process_group()
"""


def test_scalers_under_a_two_rank_nccl_group(built, tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "nccl_worker.py"
    script.write_text(_NCCL_WORKER.format(root=ROOT, port=port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "NCCL_SCALER_OK" in outs[0]


def test_extract_multi_gpu_host_gather(built):
    """One process, one thread + plan per device, every output written into slices of shared host arrays.
    On a multi-GPU box the devices are distinct and the result must be bit-equal to the 1-GPU result;
    non-default keywords (pad_to, pooled + chroma, fixed_frames) size the shared arrays."""
    import torch

    hl = built
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0]
    y = hl.synth.synth_batch(13, 30000, seed=31)
    kw = dict(n_mfcc=20, ref=np.max)
    one = hl.FeatureExtractor(device=0, **kw).extract_host(y)
    many = hl.sharding.extract_multi_gpu(y, devices, kw)
    for k in ("logmel", "mfcc", "stats", "status"):
        assert np.array_equal(one[k], many[k]), k
    opts = dict(pad_to=40000, pooled=True, chroma=True, fixed_frames=96, logmel=False)
    one = hl.FeatureExtractor(device=0, **kw).extract_host(y, **opts)
    many = hl.sharding.extract_multi_gpu(y, devices, kw, **opts)
    assert set(many) == set(one) and "logmel" not in many
    assert many["pooled"].shape == (13, 2 * 128 + 2 * 20 + 10 + 24) and many["fixed_logmel"].shape == (13, 128, 96)
    for k in one:
        assert np.array_equal(one[k], many[k], equal_nan=True), k


def test_caller_buffers_are_never_replaced(built):
    import torch

    hl = built
    ex = hl.FeatureExtractor(n_mfcc=13, ref=np.max)
    y = hl.synth.synth_batch(3, 9000, seed=2)
    T = ex.num_frames(9000)
    good = {"logmel": np.empty((3, 128, T), np.float32)}
    r = ex.extract_host(y, out=good)
    assert r["logmel"] is good["logmel"]
    for bad in (np.empty((3, 128, T + 1), np.float32), np.empty((3, 128, T), np.float64),
                np.empty((3, 128, 2 * T), np.float32)[:, :, ::2]):
        with pytest.raises(hl.ParameterError):
            ex.extract_host(y, out={"logmel": bad})
    d = torch.from_numpy(y).cuda()
    with pytest.raises(hl.ParameterError):
        ex.extract_device(d, out={"mfcc": torch.empty((3, 13, T + 2), device="cuda")})
    with pytest.raises(hl.ParameterError):
        ex.extract_host(y, pad_to=100)
    ex.close()


def test_large_clip_counts_do_not_overflow_the_grid(built):
    """per-clip helper kernels used gridDim.y = B (cap 65535)."""
    import torch

    hl = built
    B = 70000
    S = torch.rand((B, 4, 8), device="cuda") + 1e-3
    out = hl.power_to_db(S, ref=np.max)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert np.allclose(o.reshape(B, -1).max(axis=1), 0.0, atol=1e-6)
    pcm = (np.random.default_rng(0).integers(-2000, 2000, (B, 64))).astype(np.int16)
    ex = hl.FeatureExtractor(n_fft=64, hop_length=16, n_mels=8, n_mfcc=0, ref=np.max)
    r = ex.extract_host(pcm, mfcc=False, stats=False)
    assert r["logmel"].shape == (B, 8, 5) and not r["status"].any()
    ex.close()
