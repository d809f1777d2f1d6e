#!/usr/bin/env python
"""Benchmark of the audio feature hot path (BASELINE.json: clips/sec, 3 s @ 22.05 kHz,
log-mel + MFCC + spectral statistics).

  python bench.py [--gpus N --steps K --warmup W]        our CUDA path
  python bench.py --impl reference [...]                 the reference's CPU path
  torchrun ... bench.py --gpus N ...                     one rank per GPU (weak scaling)

One "step" = one pass of the full feature set over the batch BASELINE.json's configs[1]
names (10,000 synthetic 3-s clips).  `value` is device-resident throughput; `e2e` is the
same metric through the reference-facing host call (pinned host buffers, H2D and D2H
inside the timed region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

SR, CLIP_SECONDS, N_MELS, N_MFCC = 22050, 3.0, 128, 40
METRIC = "clips_per_sec_3s_22050Hz_logmel_mfcc_stats"
UNIT = "clips/s"


def rank_info():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def algorithmic_bytes_per_clip(n, T, n_mels=N_MELS, n_mfcc=N_MFCC):
    # 4n + 4T(n_mels + n_mfcc + 5)
    return 4 * n + 4 * T * (n_mels + n_mfcc + 5)


def algorithmic_flops_per_frame(n_fft=2048, n_mels=N_MELS, n_mfcc=N_MFCC, nnz=2018):
    # SURVEY.md 8(d)
    N, F = n_fft, n_fft // 2 + 1
    return N + 2.5 * N * np.log2(N) + 3 * F + F + 2 * nnz + 9 * F + N + 2 * N + 3 * n_mels + 2 * n_mfcc * n_mels


# ---------------------------------------------------------------------------
# CPU reference arm: the reference's own call pattern (five STFT-bearing librosa calls +
# zcr + rms per clip, [R] src/1_preprocessing.py:105-124) restated by the oracle.
# ---------------------------------------------------------------------------
def _cpu_one(y):
    from oracle import librosa_oracle as orc

    f = orc.extract_all_features(y, SR, with_chroma=False)
    return float(f[0])


def cpu_reference_run(clips, cores):
    """Time the oracle over `clips` (B, n) with `cores` worker processes; returns seconds."""
    from joblib import Parallel, delayed

    if cores <= 1:
        t0 = time.perf_counter()
        for y in clips:
            _cpu_one(y)
        return time.perf_counter() - t0
    with Parallel(n_jobs=cores, backend="loky") as par:
        par(delayed(_cpu_one)(clips[i]) for i in range(min(len(clips), cores)))   # spawn + import warm-up
        t0 = time.perf_counter()
        par(delayed(_cpu_one)(y) for y in clips)
        return time.perf_counter() - t0


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


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


def load_traffic():
    """DRAM bytes per launch of the frames kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "frames_fast_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=10000, help="clips per GPU per step (configs[1]: 10000)")
    ap.add_argument("--seconds", type=float, default=CLIP_SECONDS)
    ap.add_argument("--cpu-sample", type=int, default=0, help="clips in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--chroma", action="store_true", help="also compute chroma_stft (second STFT pass) in every step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank, local_rank, world = rank_info()
    n = int(round(args.seconds * SR))
    workload = (f"{args.clips}x{args.seconds:g}s@{SR}Hz synthetic white-noise clips per GPU, full set: "
                f"log-mel(n_mels={N_MELS}, ref=max, top_db=80) + MFCC({N_MFCC}) + centroid/bandwidth/rolloff/zcr/rms; "
                "n_fft=2048 hop=512 hann center pad_mode=constant")

    from hybrid_language_music_clustering_vae_b200 import synth

    if args.impl == "reference":
        if rank != 0:
            return
        for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ.setdefault(k, "1")
        cores = host_cores()
        # bounded sample: a pilot sizes each step to ~12 s / steps of CPU work on this box
        pilot = synth.synth_batch(4 * cores, n, seed=20261, mixture=False)
        rate = len(pilot) / cpu_reference_run(pilot, cores)
        per_step = args.cpu_sample or int(max(cores, min(args.clips, rate * 12.0 / max(args.steps, 1))))
        clips = synth.synth_batch(per_step, n, seed=20261, mixture=False)
        for _ in range(max(1, min(args.warmup, 1))):
            cpu_reference_run(clips[: max(cores, per_step // 4)], cores)
        t = 0.0
        for _ in range(args.steps):
            t += cpu_reference_run(clips, cores)
        v = per_step * args.steps / t
        sample = (f"{per_step} clips per step x {args.steps} steps of the same workload; oracle port of the "
                  f"reference's extract_all_features (5 redundant STFTs, no chroma) under joblib n_jobs={cores}")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload, "clips_per_step": per_step},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return

    import torch
    import torch.distributed as dist
    import hybrid_language_music_clustering_vae_b200 as hl

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

    ex = hl.FeatureExtractor(sr=SR, n_fft=2048, hop_length=512, n_mels=N_MELS, n_mfcc=N_MFCC, ref=np.max,
                             device=local_rank)
    assert ex.uses_fast_path()
    T = ex.num_frames(n)
    B = args.clips

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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        out = ex.extract_device(d_wave, out=out, chroma=args.chroma)
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = hl.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    frames_ms, db_ms, calls = ex.read_timing()
    ex.set_timing(False)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)
    status_bad = int(out["status"].sum().item())

    # ---- e2e: pinned host in, pinned host out, through the C ABI's host pipeline
    e2e = None
    if not args.no_e2e:
        h_out = {
            "logmel": torch.empty((B, N_MELS, T), dtype=torch.float32, pin_memory=True).numpy(),
            "mfcc": torch.empty((B, N_MFCC, T), dtype=torch.float32, pin_memory=True).numpy(),
            "stats": torch.empty((B, 5, T), dtype=torch.float32, pin_memory=True).numpy(),
            "status": torch.empty((B,), dtype=torch.int32, pin_memory=True).numpy(),
        }
        hw = h_wave_t.numpy()
        for _ in range(2):
            ex.extract_host(hw, out=h_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ex.extract_host(hw, out=h_out)      # blocking: returns when the last D2H has landed
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        h2d, d2h = ex.last_transfer_bytes()
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # the bound of this path: a bare pinned-host -> device copy of the same input (SURVEY 8d asks for it next
        # to the result); measured on this rank with nothing else on the bus
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d_flat = d_store.view(-1)[: B * n].view(B, n)
        d_flat.copy_(h_wave_t, non_blocking=True)
        torch.cuda.synchronize(dev)
        c0.record()
        for _ in range(3):
            d_flat.copy_(h_wave_t, non_blocking=True)
        c1.record()
        torch.cuda.synchronize(dev)
        h2d_peak = 3 * h_wave_t.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        d_wave.copy_(h_wave_t, non_blocking=True)            # restore the pitched device copy
        torch.cuda.synchronize(dev)
        e2e = {"value": world * B * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * dt / args.steps,
               "pcie_gbs": (h2d + d2h) * args.steps / dt / 1e9,
               "h2d_gbs": h2d * args.steps / dt / 1e9, "h2d_bare_copy_gbs": h2d_peak,
               "bound": "PCIe host->device: the step's input alone takes h2d_bytes / h2d_bare_copy_gbs",
               "checksum": float(h_out["logmel"][0, 0, :4].sum())}

    # ---- secondary e2e: the same clips as 16-bit PCM (what the WAV files hold); the device does
    #      librosa.load's int16/32768 conversion, so H2D bytes halve.  Reported, not the headline.
    e2e_pcm = None
    if not args.no_e2e:
        h_pcm_t = torch.empty((B, n), dtype=torch.int16, pin_memory=True)
        h_pcm = h_pcm_t.numpy()
        hw_f = h_wave_t.numpy()
        for lo in range(0, B, 512):
            h_pcm[lo:lo + 512] = np.clip(np.rint(hw_f[lo:lo + 512] * 32768.0), -32768, 32767).astype(np.int16)
        for _ in range(2):
            ex.extract_host(h_pcm, out=h_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ex.extract_host(h_pcm, out=h_out)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        h2d, d2h = ex.last_transfer_bytes()
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e_pcm = {"value": world * B * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * dt / args.steps,
                   "input": "int16 PCM host buffers, converted on the device (librosa.load semantics)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (frames_fast_2048), timed by events on its stream
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    k1_ms = frames_ms / max(calls, 1)
    k1_bytes = B * (4 * n + 4 * T * (N_MELS + 5))                  # what the frames kernel itself must move
    step_bytes = B * algorithmic_bytes_per_clip(n, T)
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
    traffic = load_traffic()
    ncu_pipe = None
    try:      # FP32-pipe / issue / shared-memory utilisation of the same kernel from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "frames_fast_pipes.json")) as f:
            ncu_pipe = json.load(f)
    except Exception:
        pass
    fp32_peak = hl.measure_fp32_peak(local_rank)
    flops_step = B * T * algorithmic_flops_per_frame()
    roofline = {
        "bound": "hbm", "kernel": "frames_fast_2048", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "peak_source": peak_src,
        "traffic": (traffic or {}).get("dram_bytes_per_launch"),
        "algorithmic_bytes_per_launch": k1_bytes, "kernel_ms": k1_ms,
        "kernel_share_of_step": frames_ms / (frames_ms + db_ms) if (frames_ms + db_ms) > 0 else None,
        "db_dct_ms": db_ms / max(calls, 1),
        "step_achieved_gbs": step_bytes / (ms / args.steps * 1e-3) / 1e9,
        "note": "the path is FP32-issue bound, not HBM bound (SURVEY 8d); see fp32",
        "fp32": {"achieved_tflops": flops_step / (ms / args.steps * 1e-3) / 1e12,
                 "peak_tflops": fp32_peak, "peak_source": "FMA micro-benchmark run in this process",
                 "frac": flops_step / (ms / args.steps * 1e-3) / 1e12 / fp32_peak,
                 "algorithmic_flops_per_frame": algorithmic_flops_per_frame(),
                 "ncu": ncu_pipe},
    }

    cpu = None
    if not args.no_cpu and world == 1:       # the CPU baseline is reported at N = 1 only
        for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ.setdefault(k, "1")
        cores = host_cores()
        pilot_n = min(B, 4 * cores)
        rate = pilot_n / cpu_reference_run(h_wave_t.numpy()[:pilot_n], cores)
        ns = args.cpu_sample or int(max(cores, min(B, rate * 12.0)))      # ~12 s of CPU work
        t = cpu_reference_run(h_wave_t.numpy()[:ns], cores)
        cpu = {"value": ns / t, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {ns} clips of the same batch, oracle port of the reference's "
                         f"extract_all_features (5 redundant STFTs, no chroma), joblib n_jobs={cores}, {t:.1f} s"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "clips_per_gpu": B, "samples_per_clip": n, "frames_per_clip": T,
                   "chroma_stft_included": bool(args.chroma),
                   "l2_policy": "inputs larger than L2 (2.6 GB per step)", "parallelism": f"clips sharded x{world}, no collective",
                   "numa_bound": bool(numa_bound)},
        "audio_hours_per_sec": value * args.seconds / 3600.0,
        "e2e": e2e, "e2e_pcm16": e2e_pcm, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu, "nonfinite_clips": status_bad,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
