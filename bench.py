#!/usr/bin/env python
"""Benchmark of the audio feature hot path (BASELINE.json: clips/sec, 3 s @ 22.05 kHz,
log-mel + MFCC + spectral statistics).

  python bench.py [--gpus N --steps K --warmup W]        our CUDA path
  python bench.py --impl reference [...]                 the reference's CPU path
  torchrun ... bench.py --gpus N ...                     one rank per GPU (weak scaling)

One "step" = one pass of the full feature set over the batch BASELINE.json's configs[1]
names (10,000 synthetic 3-s clips per GPU).  `value` is device-resident throughput; `e2e`
is the same metric through the reference-facing host call (pinned host buffers, H2D and
D2H inside the timed region).  Prints ONE JSON line on rank 0.

The reference arm never imports the product package (it would map libhlmc_b200.so): the
synthetic-clip generator is loaded by file path, the arithmetic is oracle/librosa_oracle.py.
"""
from __future__ import annotations

import argparse
import hashlib
import importlib.util
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "hybrid_language_music_clustering_vae_b200")

import numpy as np

SR, CLIP_SECONDS, N_MELS, N_MFCC = 22050, 3.0, 128, 40
UNIT = "clips/s"


def metric_name(seconds):
    return f"clips_per_sec_{seconds:g}s_22050Hz_logmel_mfcc_stats"
L2_BYTES = 126 << 20


def rank_info():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def load_synth():
    """hybrid_language_music_clustering_vae_b200/synth.py by path: pure numpy, no package import, so the
    reference arm's process never maps the CUDA library."""
    spec = importlib.util.spec_from_file_location("hlmc_synth_by_path", os.path.join(PKG, "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def algorithmic_bytes_per_clip(n, T, n_mels=N_MELS, n_mfcc=N_MFCC):
    # 4n + 4T(n_mels + n_mfcc + 5)
    return 4 * n + 4 * T * (n_mels + n_mfcc + 5)


def algorithmic_flops_per_frame(n_fft=2048, n_mels=N_MELS, n_mfcc=N_MFCC, nnz=2018):
    # SURVEY.md 8(d)
    N, F = n_fft, n_fft // 2 + 1
    return N + 2.5 * N * np.log2(N) + 3 * F + F + 2 * nnz + 9 * F + N + 2 * N + 3 * n_mels + 2 * n_mfcc * n_mels


def make_config(args, world):
    """The `config` object: identical keys and values in both arms (the driver compares them)."""
    n = int(round(args.seconds * SR))
    T = 1 + n // 512
    step_bytes = args.clips * algorithmic_bytes_per_clip(n, T)
    return {
        "workload": (f"{args.clips}x{args.seconds:g}s@{SR}Hz synthetic white-noise clips per GPU, full set: "
                     f"log-mel(n_mels={N_MELS}, ref=max, top_db=80) + MFCC({N_MFCC}) + "
                     "centroid/bandwidth/rolloff/zcr/rms; n_fft=2048 hop=512 hann center pad_mode=constant"),
        "clips_per_gpu": args.clips, "samples_per_clip": n, "frames_per_clip": T,
        "chroma_stft_included": bool(args.chroma),
        "l2_policy": ("inputs larger than L2" if step_bytes > 2 * L2_BYTES else "L2 flushed between steps")
                     + f" ({step_bytes / 1e9:.3f} GB per step per GPU, L2 0.132 GB)",
        "parallelism": f"clips sharded x{world}, no collective",
    }


# ---------------------------------------------------------------------------
# CPU reference arm: the reference's own call pattern (five STFT-bearing librosa calls +
# zcr + rms per clip, [R] src/1_preprocessing.py:105-124) restated by the oracle.  Workers map the
# clips from one .npy file (no per-task pickling of audio) and take blocks of clips, so that the
# measurement is the arithmetic, not joblib's dispatch.
# ---------------------------------------------------------------------------
class CpuArm:
    """joblib pool over all host cores + the clips in a memory-mapped file."""

    def __init__(self, clips, cores):
        from joblib import Parallel

        self.cores = cores
        self.dir = tempfile.mkdtemp(prefix="hlmc_cpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        self.path = os.path.join(self.dir, "clips.npy")
        np.save(self.path, clips)
        self.n = len(clips)
        self.par = Parallel(n_jobs=cores, backend="loky") if cores > 1 else None
        if self.par is not None:
            self.par.__enter__()
            self.run(min(self.n, 2 * cores))          # spawn the workers, import scipy, map the file

    def run(self, count=None, block=8):
        """Process the first `count` clips; returns seconds."""
        from joblib import delayed

        from oracle import bench_worker

        count = self.n if count is None else min(count, self.n)
        t0 = time.perf_counter()
        if self.par is None:
            bench_worker.block(self.path, 0, count)
        else:
            self.par(delayed(bench_worker.block)(self.path, lo, min(count, lo + block))
                     for lo in range(0, count, block))
        return time.perf_counter() - t0

    def close(self):
        if self.par is not None:
            self.par.__exit__(None, None, None)
        try:
            os.remove(self.path)
            os.rmdir(self.dir)
        except OSError:
            pass


def single_thread_rate(clips, count=48):
    """1 process, 1 thread: how [R] src/1_preprocessing.py:232 runs its per-file loop."""
    from oracle import librosa_oracle as orc

    count = min(count, len(clips))
    orc.extract_all_features(clips[0], SR, with_chroma=False)
    t0 = time.perf_counter()
    for i in range(count):
        orc.extract_all_features(clips[i], SR, with_chroma=False)
    return count / (time.perf_counter() - t0), count


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


CPU_SAMPLE_TEXT = ("oracle port of the reference's extract_all_features (5 redundant STFTs + zcr + rms, no chroma) "
                   "under joblib n_jobs={cores}, blocks of 8 clips per task, clips memory-mapped")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def kernel_source_hash():
    """sha256 over the sources of the captured kernels (hlmc_kernels.cu and the headers it includes): an ncu
    capture is only quoted when it was taken from these exact files."""
    h = hashlib.sha256()
    csrc = os.path.join(PKG, "csrc")
    for name in ("fft_inreg.cuh", "fft_inreg2.cuh", "hlmc_internal.h", "hlmc_kernels.cu"):
        with open(os.path.join(csrc, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def load_capture(name):
    """A committed ncu summary (profiles/<name>), refused unless its `source_hash` matches the built sources."""
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            d = json.load(f)
    except Exception:
        return None, "no capture committed"
    if d.get("source_hash") != kernel_source_hash():
        return None, f"stale capture (taken from sources {d.get('source_hash')}, built sources {kernel_source_hash()})"
    return d, "ncu capture of these sources"


def pcie_links():
    """PCIe generation / width of every GPU, from nvidia-smi."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=index,pcie.link.gen.current,pcie.link.gen.max,"
                              "pcie.link.width.current,pci.bus_id", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=20).stdout
        links = []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            links.append({"gpu": int(f[0]), "gen": f[1], "gen_max": f[2], "width": f[3], "bus": f[4]})
        return links
    except Exception as e:  # pragma: no cover
        return [{"error": str(e)}]


def reference_arm(args, rank, world):
    if rank != 0:
        return
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(k, "1")
    synth = load_synth()
    n = int(round(args.seconds * SR))
    cores = host_cores()
    pilot = synth.synth_batch(max(128, 32 * cores), n, seed=20261, mixture=False)
    arm = CpuArm(pilot, cores)
    rate = len(pilot) / arm.run()
    arm.close()
    # bounded sample, but long steps: >= 4,000 clips per step (pool dispatch and the straggler tail are then
    # < 2 % of a step) unless that would push the whole run past ~6 minutes
    per_step = int(min(args.clips, max(4000, rate * 170.0 / (args.steps + 1))))
    if per_step * (args.steps + 1) / rate > 360.0:
        per_step = int(max(8 * cores, rate * 360.0 / (args.steps + 1)))
    per_step = args.cpu_sample or per_step
    clips = synth.synth_batch(per_step, n, seed=20261, mixture=False)
    st_rate, st_n = single_thread_rate(clips)
    arm = CpuArm(clips, cores)
    for _ in range(1 if args.warmup > 0 else 0):
        arm.run(max(8 * cores, per_step // 8))
    t = 0.0
    for _ in range(args.steps):
        t += arm.run()
    arm.close()
    v = per_step * args.steps / t
    sample = (f"{per_step} clips per step x {args.steps} steps of the same workload; " +
              CPU_SAMPLE_TEXT.format(cores=cores))
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args.seconds), "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": make_config(args, world),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "cpu_model": cpu_model(),
                         "single_thread": {"value": st_rate, "unit": UNIT, "cores": 1,
                                           "sample": f"{st_n} clips, 1 process 1 thread ([R] src/1_preprocessing.py:232)"}},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "clips_per_step": per_step,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=10000, help="clips per GPU per step (configs[1]: 10000)")
    ap.add_argument("--seconds", type=float, default=CLIP_SECONDS)
    ap.add_argument("--cpu-sample", type=int, default=0, help="clips in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--chroma", action="store_true", help="also compute chroma_stft in every device-resident step")
    ap.add_argument("--e2e-mode", default="full_f32",
                    choices=["full_f32", "full_pcm16", "basic_contract", "basic_contract_pcm16", "advanced_contract"],
                    help="which end-to-end mode is the headline `e2e` (all are reported under e2e_modes)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-probes", action="store_true", help="skip the host-memory / concurrent-H2D probes")
    args = ap.parse_args()
    rank, local_rank, world = rank_info()
    n = int(round(args.seconds * SR))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import hybrid_language_music_clustering_vae_b200 as hl
    from hybrid_language_music_clustering_vae_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one rank per GPU: keep the rank's pinned buffers on the GPU's own NUMA node
    numa_bound = hl.sharding.bind_to_gpu_numa_node(local_rank) if world > 1 else False
    if world > 1:
        # NCCL prints its version banner on the C-level stdout at the first collective; keep
        # stdout clean for the single JSON line by pointing fd 1 at stderr while it initialises
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x, dtype=torch.float64):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=dtype)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(x):
        return -max_over_ranks(-x)

    ex = hl.FeatureExtractor(sr=SR, n_fft=2048, hop_length=512, n_mels=N_MELS, n_mfcc=N_MFCC, ref=np.max,
                             device=local_rank)
    assert ex.uses_fast_path()
    T = ex.num_frames(n)
    B = args.clips
    config = make_config(args, world)
    step_bytes = B * algorithmic_bytes_per_clip(n, T)
    flush = None
    if step_bytes <= 2 * L2_BYTES:            # small batches: evict L2 between timed steps
        flush = torch.empty((3 * L2_BYTES,), dtype=torch.uint8, device=dev)

    # synthetic inputs: pinned host copy (for e2e) and a pitched device copy (for `value`)
    pitch = (n + 3) & ~3
    h_wave_t = torch.empty((B, n), dtype=torch.float32, pin_memory=True)
    synth.synth_batch(B, n, seed=20261 + rank, mixture=False, out=h_wave_t.numpy())
    d_store = torch.empty((B, pitch), dtype=torch.float32, device=dev)
    d_wave = d_store[:, :n]
    d_wave.copy_(h_wave_t, non_blocking=True)
    torch.cuda.synchronize(dev)

    out = None
    for _ in range(max(3, args.warmup)):
        out = ex.extract_device(d_wave, out=out, chroma=args.chroma)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ex.set_timing(True)
    ex.read_timing()
    launches0 = hl.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = ex.extract_device(d_wave, out=out, chroma=args.chroma)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
    else:
        evs = []
        for _ in range(args.steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = ex.extract_device(d_wave, out=out, chroma=args.chroma)
            b.record()
            evs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
    t_wall1 = time.perf_counter()
    launches = hl.launch_count() - launches0
    frames_ms, db_ms, calls = ex.read_timing()
    ex.set_timing(False)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms = max_over_ranks(ms)
    value = world * B * args.steps / (ms * 1e-3)
    status_bad = int(out["status"].sum().item())
    del out
    torch.cuda.empty_cache()

    # ---- end to end: pinned host in, pinned host out, through the C ABI's host pipeline (hlmc_extract_host_io).
    #      Three contracts: full arrays out; what 1_preprocessing.py keeps (370 pooled
    #      columns incl. chroma); what 1_preprocessing_advanced.py keeps ((128,1024) image + 290 columns).
    e2e_modes, probes = {}, {}
    if not args.no_e2e:
        pin = lambda shape, dt=torch.float32: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
        hw = h_wave_t.numpy()
        h_pcm = None

        def pcm():
            nonlocal h_pcm
            if h_pcm is None:
                h_pcm = pin((B, n), torch.int16)
                for lo in range(0, B, 512):
                    h_pcm[lo:lo + 512] = np.clip(np.rint(hw[lo:lo + 512] * 32768.0), -32768, 32767).astype(np.int16)
            return h_pcm

        def run_mode(name, desc, call, checksum):
            for _ in range(2):
                call()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                call()                       # blocking: returns when the last D2H has landed
            torch.cuda.synchronize(dev)
            dt = max_over_ranks(time.perf_counter() - t0)
            h2d, d2h = ex_of[name].last_transfer_bytes()
            e2e_modes[name] = {
                "value": world * B * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * dt / args.steps,
                "h2d_gbs_per_gpu": h2d * args.steps / dt / 1e9, "d2h_gbs_per_gpu": d2h * args.steps / dt / 1e9,
                "contract": desc, "checksum": float(checksum())}

        ex_of = {}
        full_out = {"logmel": pin((B, N_MELS, T)), "mfcc": pin((B, N_MFCC, T)), "stats": pin((B, 5, T)),
                    "status": pin((B,), torch.int32)}
        ex_of["full_f32"] = ex_of["full_pcm16"] = ex
        run_mode("full_f32", "float32 clips in -> log-mel (B,128,T) + MFCC (B,40,T) + stats (B,5,T) out",
                 lambda: ex.extract_host(hw, out=full_out), lambda: full_out["logmel"][0, 0, :4].sum())
        run_mode("full_pcm16", "int16 PCM clips in (converted on the device as librosa.load does) -> the same arrays out",
                 lambda: ex.extract_host(pcm(), out=full_out), lambda: full_out["logmel"][0, 0, :4].sum())
        del full_out
        # [R] src/1_preprocessing.py:105-129 keeps 370 floats per clip (pooled columns incl. chroma_stft)
        basic_out = {"pooled": pin((B, ex.pooled_width(True, True))), "status": pin((B,), torch.int32)}
        ex_of["basic_contract"] = ex_of["basic_contract_pcm16"] = ex
        basic = lambda w: ex.extract_host(w, logmel=False, mfcc=False, stats=False, pooled=True, chroma="pooled",
                                          out=basic_out)
        run_mode("basic_contract", "1_preprocessing.py: float32 clips in -> (B,370) pooled columns incl. chroma_stft out",
                 lambda: basic(hw), lambda: basic_out["pooled"][0, :4].sum())
        run_mode("basic_contract_pcm16", "1_preprocessing.py from int16 PCM: -> (B,370) pooled columns incl. chroma_stft",
                 lambda: basic(pcm()), lambda: basic_out["pooled"][0, :4].sum())
        # [R] src/1_preprocessing_advanced.py:97-156 keeps a (128, 1024) image + 290 floats per clip
        ex_adv = hl.FeatureExtractor(sr=SR, n_fft=2048, hop_length=512, n_mels=N_MELS, n_mfcc=0, ref=np.max,
                                     device=local_rank)
        fixed = 1024
        adv_out = {"fixed_logmel": pin((B, N_MELS, fixed)), "pooled": pin((B, ex_adv.pooled_width(False, True))),
                   "status": pin((B,), torch.int32)}
        ex_of["advanced_contract"] = ex_adv
        run_mode("advanced_contract",
                 "1_preprocessing_advanced.py: float32 clips in -> (B,128,1024) padded log-mel image + (B,290) columns out",
                 lambda: ex_adv.extract_host(hw, logmel=False, mfcc=False, stats=False, pooled=True, chroma="pooled",
                                             fixed_frames=fixed, out=adv_out),
                 lambda: adv_out["fixed_logmel"][0, 0, :4].sum())
        del adv_out
        ex_adv.close()

        if not args.no_probes:
            # what bounds the host paths: (1) this rank's bare pinned H2D copy, alone on the bus; (2) the same with
            # every rank copying at once; (3) a pinned -> pinned host memcpy on every rank at once (host DRAM)
            def h2d_gbs(reps=3):
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                d_flat = d_store.view(-1)[: B * n].view(B, n)
                d_flat.copy_(h_wave_t, non_blocking=True)
                torch.cuda.synchronize(dev)
                c0.record()
                for _ in range(reps):
                    d_flat.copy_(h_wave_t, non_blocking=True)
                c1.record()
                torch.cuda.synchronize(dev)
                return reps * h_wave_t.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9

            alone = []
            for r in range(world):          # one rank at a time
                barrier()
                if r == rank:
                    alone.append(h2d_gbs())
                barrier()
            barrier()
            together = h2d_gbs()
            barrier()
            h_dst = torch.empty_like(h_wave_t).pin_memory() if B * n * 4 <= (4 << 30) else None
            memcpy_gbs = None
            if h_dst is not None:
                h_dst.copy_(h_wave_t)
                barrier()
                t0 = time.perf_counter()
                for _ in range(2):
                    h_dst.copy_(h_wave_t)
                memcpy_gbs = 2 * 2 * h_wave_t.numel() * 4 / (time.perf_counter() - t0) / 1e9     # read + write
                del h_dst
            probes = {"h2d_alone_gbs_this_rank": alone[0] if alone else None,
                      "h2d_all_ranks_concurrent_gbs_min": min_over_ranks(together),
                      "h2d_all_ranks_concurrent_gbs_max": max_over_ranks(together),
                      "host_memcpy_all_ranks_concurrent_gbs_min": (min_over_ranks(memcpy_gbs) if memcpy_gbs else None),
                      "host_memcpy_note": "pinned->pinned torch copy_ (read+write bytes), one thread per rank",
                      "numa_bound": bool(numa_bound)}
            d_wave.copy_(h_wave_t, non_blocking=True)            # restore the pitched device copy
            torch.cuda.synchronize(dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if probes:
        probes["pcie_links"] = pcie_links()
        probes["host_cores"] = host_cores()

    # ---- roofline of the dominant kernel (the frames kernel), timed by events on its stream inside the C ABI.
    #      The path is FP32-pipe bound (34 flop/B against a ridge of ~11), so that is `bound`;
    #      the HBM view is kept alongside.
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "of measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "of fallback 6650 GB/s"
    k1_ms = frames_ms / max(calls, 1)
    k2_ms = db_ms / max(calls, 1)
    k1_bytes = B * (4 * n + 4 * T * (N_MELS + 5))                  # what the frames kernel itself must move
    fp32_peak = hl.measure_fp32_peak(local_rank)
    # per frame: everything of SURVEY 8(d)'s count except log10+clamp and the DCT, which run in db_dct
    k1_flops = B * T * (algorithmic_flops_per_frame() - 3 * N_MELS - 2 * N_MFCC * N_MELS)
    step_flops = B * T * algorithmic_flops_per_frame()
    k1_tflops = k1_flops / (k1_ms * 1e-3) / 1e12
    traffic, traffic_note = load_capture("frames_fast_traffic.json")
    pipes, pipes_note = load_capture("frames_fast_pipes.json")
    step_ms = ms / args.steps
    roofline = {
        "bound": "fp32", "kernel": "frames_fast_2048<16 warps, TM> (per-lane tables in Tensor Memory, tcgen05.ld)",
        "achieved": k1_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": k1_tflops / fp32_peak,
        "peak_source": "FP32 FMA micro-benchmark run in this process (MEASURED_PEAKS.json has no fp32 entry; "
                       "nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5)",
        "algorithmic_flops_per_launch": k1_flops, "kernel_ms": k1_ms,
        "kernel_share_of_step": frames_ms / (frames_ms + db_ms) if (frames_ms + db_ms) > 0 else None,
        "db_dct_ms": k2_ms,
        "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_source": traffic_note,
        "ncu_pipes": pipes, "ncu_pipes_source": pipes_note,
        "step": {"fp32_tflops": step_flops / (step_ms * 1e-3) / 1e12,
                 "fp32_frac": step_flops / (step_ms * 1e-3) / 1e12 / fp32_peak,
                 "hbm_gbs": step_bytes / (step_ms * 1e-3) / 1e9,
                 "hbm_frac": step_bytes / (step_ms * 1e-3) / 1e9 / hbm_peak,
                 "algorithmic_flops_per_frame": algorithmic_flops_per_frame(),
                 "algorithmic_bytes_per_clip": algorithmic_bytes_per_clip(n, T)},
        "hbm": {"achieved": k1_bytes / (k1_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": k1_bytes / (k1_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                "algorithmic_bytes_per_launch": k1_bytes},
    }

    cpu = None
    if not args.no_cpu and world == 1:       # the CPU baseline is reported at N = 1 only
        for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ.setdefault(k, "1")
        cores = host_cores()
        hw = h_wave_t.numpy()
        arm = CpuArm(hw[: min(B, max(128, 32 * cores))], cores)
        rate = arm.n / arm.run()
        arm.close()
        # one long step: >= 4,000 clips (what the reference arm uses per step), ~15-30 s of CPU work
        ns = args.cpu_sample or int(min(B, max(4000, rate * 15.0)))
        arm = CpuArm(hw[:ns], cores)
        t = arm.run()
        arm.close()
        st_rate, st_n = single_thread_rate(hw)
        cpu = {"value": ns / t, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": cpu_model(),
               "sample": f"first {ns} clips of the same batch in one step of {t:.1f} s; " + CPU_SAMPLE_TEXT.format(cores=cores),
               "single_thread": {"value": st_rate, "unit": UNIT, "cores": 1,
                                 "sample": f"{st_n} clips, 1 process 1 thread ([R] src/1_preprocessing.py:232)"}}

    line = {
        "metric": metric_name(args.seconds), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "audio_hours_per_sec": value * args.seconds / 3600.0,
        "e2e": e2e_modes.get(args.e2e_mode), "e2e_mode": args.e2e_mode if e2e_modes else None,
        "e2e_modes": e2e_modes or None, "host_probes": probes or None,
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu, "nonfinite_clips": status_bad,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
