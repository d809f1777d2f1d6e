"""The C-ABI boundary and the host logic, without a GPU."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "hlmc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hlmc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from hybrid_language_music_clustering_vae_b200 import _lib

    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} is declared in include/hlmc_b200.h but not exported"
    assert set(_lib.EXPORTS) == set(names)
    assert _lib.lib.hlmc_abi_version() == 2 == _lib.ABI_VERSION


def test_params_struct_matches_header(built):
    from hybrid_language_music_clustering_vae_b200._lib import HlmcParams, lib

    src = open(os.path.join(ROOT, "include", "hlmc_b200.h")).read()
    body = src[src.index("typedef struct hlmc_params {"):src.index("} hlmc_params;")]
    fields = re.findall(r"^\s*(int32_t|float)\s+(\w+);", body, flags=re.M)
    assert [f for _t, f in fields] == [f for f, _t in HlmcParams._fields_]
    assert C.sizeof(HlmcParams) == 4 * len(fields)
    p = HlmcParams()
    lib.hlmc_params_default(C.byref(p))
    assert (p.sr, p.n_fft, p.hop_length, p.n_mels, p.n_mfcc) == (22050, 2048, 512, 128, 20)
    assert p.pad_mode == 0 and p.center == 1 and abs(p.top_db - 80.0) < 1e-6 and abs(p.amin - 1e-10) < 1e-16


def test_host_io_struct_matches_header(built):
    """Field order and types of hlmc_host_io (ABI 2) as the ctypes mirror declares them."""
    from hybrid_language_music_clustering_vae_b200._lib import HlmcHostIo

    src = open(os.path.join(ROOT, "include", "hlmc_b200.h")).read()
    body = src[src.index("typedef struct hlmc_host_io {"):src.index("} hlmc_host_io;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in re.findall(r"^\s*(?:const\s+)?(?:void|float|int32_t|int64_t)\s*\**\s*([\w\s,\*]+);", body, flags=re.M):
        names += [n.strip().lstrip("*") for n in decl.split(",")]
    assert names == [f for f, _t in HlmcHostIo._fields_]
    assert C.sizeof(HlmcHostIo) % 8 == 0


def test_reference_arm_never_maps_the_product_library():
    """bench.py --impl reference must not import the package (the arm's record listed libhlmc_b200.so)."""
    code = ("import sys, runpy; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0','--cpu-sample','8'];"
            "runpy.run_path(%r, run_name='__main__');"
            "assert not any('hybrid_language_music_clustering_vae_b200' == m.split('.')[0] for m in sys.modules), 'package imported';"
            "assert 'libhlmc_b200' not in open('/proc/self/maps').read(), 'library mapped'; print('CLEAN')" % os.path.join(ROOT, "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "CLEAN" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("n,T", [(66150, 130), (661500, 1292), (2047, 4), (2048, 5), (2049, 5), (511, 1)])
def test_num_frames_matches_oracle(built, n, T):
    from hybrid_language_music_clustering_vae_b200._lib import HlmcParams, lib
    from oracle import librosa_oracle as orc

    p = HlmcParams()
    lib.hlmc_params_default(C.byref(p))
    assert lib.hlmc_num_frames(C.byref(p), n) == T == orc.num_frames(n)
    p.center = 0
    if n >= 2048:
        assert lib.hlmc_num_frames(C.byref(p), n) == orc.num_frames(n, center=False)
    else:
        assert lib.hlmc_num_frames(C.byref(p), n) < 0      # librosa raises ParameterError


def test_no_cpu_fallback(built):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch

    hl = built
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hl.FeatureExtractor()
    with pytest.raises(RuntimeError):
        hl.feature.melspectrogram(y=np.zeros(4096, np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hybrid_language_music_clustering_vae_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                for banned in ("cufft", "torchaudio", "torch.stft", "import triton", "import librosa"):
                    assert banned not in txt, (f, banned)


def test_python_side_parameter_errors(built):
    hl = built
    with pytest.raises(hl.ParameterError):
        hl.FeatureExtractor(pad_mode="wrap")
    with pytest.raises(hl.ParameterError):
        hl.FeatureExtractor(top_db=-1.0)
    with pytest.raises(hl.ParameterError):
        hl.power_to_db(np.ones((2, 2), np.float32), amin=0.0)
    with pytest.raises(hl.ParameterError):
        hl.feature.spectral_rolloff(y=np.zeros(4096, np.float32), roll_percent=1.5)
    with pytest.raises(hl.ParameterError):
        hl.feature.melspectrogram(y=np.zeros(16, np.int16))
    with pytest.raises(hl.UnsupportedError):
        hl.feature.mfcc(y=np.zeros(4096, np.float32), dct_type=3)


def test_shard_bounds_cover_the_batch(built):
    from hybrid_language_music_clustering_vae_b200.sharding import shard_bounds

    for B in (0, 1, 7, 8, 1000, 100001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (l0, h0), (l1, h1) in zip(spans, spans[1:]):
                assert h0 == l1 and l0 <= h0
            assert max(h - l for l, h in spans) <= -(-B // world) if B else True


def test_synth_is_deterministic_and_in_range(built):
    hl = built
    a = hl.synth.synth_batch(20, 3000, seed=5)
    b = hl.synth.synth_batch(20, 3000, seed=5)
    assert np.array_equal(a, b) and a.dtype == np.float32
    assert np.abs(a).max() <= 1.0
    kinds = hl.synth.mixture_kinds(1000)
    frac = {k: kinds.count(k) / 1000 for k in set(kinds)}
    assert abs(frac["white"] - 0.40) < 0.02 and abs(frac["harmonic"] - 0.30) < 0.02
    assert np.all(a[kinds[:20].index("halfsilent")][1500:] == 0) if "halfsilent" in kinds[:20] else True


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from hybrid_language_music_clustering_vae_b200.sharding import shard_bounds, gather_host
from oracle import librosa_oracle as orc
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
B, n = 5, 4096
rng = np.random.default_rng(0)
y = (0.1 * rng.standard_normal((B, n))).astype(np.float32)     # same batch on every rank
lo, hi = shard_bounds(B, rank, world)
# stand-in for the per-rank GPU extraction: the CPU oracle (allowed in tests only)
local = np.stack([orc.rms(y=y[i])[0] for i in range(lo, hi)]) if hi > lo else np.zeros((0, 9), np.float32)
full = gather_host(local, B, rank, world)
if rank == 0:
    want = np.stack([orc.rms(y=y[i])[0] for i in range(B)])
    assert full.shape == want.shape and np.array_equal(full, want)
    print("GATHER_OK")
else:
    assert full is None
dist.destroy_process_group()
"""


def test_two_rank_gloo_host_gather(tmp_path):
    """World-size-2 run of the sharding + host gather path on CPU (gloo), no collective on the data path."""
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, port=port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-sample", "8"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "clips/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config"):
        assert k in line
    st = line["cpu_baseline"]["single_thread"]
    assert st["cores"] == 1 and st["value"] > 0
    # both arms build `config` with the same function, so the driver's same_config check can hold
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    import argparse
    args = argparse.Namespace(clips=10000, seconds=3.0, chroma=False)
    assert line["config"] == bench.make_config(args, 1)


def test_chan_combination_of_shard_statistics(built):
    """The cross-GPU step of the device StandardScaler, on the host: shard statistics combine exactly."""
    from hybrid_language_music_clustering_vae_b200.scaler import combine_stats, make_sklearn_scaler
    from sklearn.preprocessing import StandardScaler

    rng = np.random.default_rng(0)
    X = rng.standard_normal((101, 33)) * 5 + 2
    shards = [X[:10], X[10:64], X[64:64], X[64:]]
    cnt = [len(s) for s in shards]
    means = [s.mean(0) if len(s) else np.zeros(33) for s in shards]
    m2s = [((s - s.mean(0)) ** 2).sum(0) if len(s) else np.zeros(33) for s in shards]
    n, mean, m2 = combine_stats(cnt, means, m2s)
    ref = StandardScaler().fit(X)
    assert n == 101 and np.allclose(mean, ref.mean_) and np.allclose(m2 / n, ref.var_)
    sc = make_sklearn_scaler(n, mean, m2 / n)
    assert np.allclose(sc.transform(X), ref.transform(X))


def test_numa_binding_helpers_are_safe_without_a_gpu(built):
    """bench.py pins each rank next to its GPU before allocating pinned buffers; without a GPU (or without
    sysfs NUMA information) the helper must change nothing and report False."""
    sh = built.sharding
    assert sh._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sh._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert sh.bind_to_gpu_numa_node(0) in (False, True)
    import torch

    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
