"""Multi-GPU sharding of a clip batch.

Clips are independent, so there is no collective on the data path: rank r takes
the contiguous block ``shard_bounds(B, r, world)``, runs it on its own GPU and
streams, and the results meet in host memory (``gather_host`` across processes,
or plain slice writes across threads in ``extract_multi_gpu``).
"""
from __future__ import annotations

import os
import threading

import numpy as np


def shard_bounds(B: int, rank: int, world: int):
    """Contiguous block [lo, hi) of rank `rank`: ceil(B/world) clips each, last ranks may be short."""
    per = -(-B // world) if world > 0 else B
    lo = min(B, rank * per)
    hi = min(B, lo + per)
    return lo, hi


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_cpus(device: int):
    """CPUs of the NUMA node the GPU `device` hangs off (sysfs), or None if unknown."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(device)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_numa_node(device: int) -> bool:
    """Pin the calling process to the CPUs next to GPU `device` BEFORE it allocates pinned host
    buffers, so that first-touch places them on the GPU's NUMA node: with 8 ranks on a two-socket
    box the H2D / D2H copies otherwise cross the socket link for half of the GPUs.  Returns True
    if the affinity was changed.  No collective, no effect on results."""
    cpus = gpu_numa_cpus(device)
    if not cpus:
        return False
    try:
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if not target or target == allowed:
            return False
        os.sched_setaffinity(0, target)
        return True
    except Exception:
        return False


def gather_host(local: np.ndarray, B: int, rank: int, world: int, group=None, dst: int = 0):
    """Host gather of per-rank slices (leading axis) onto rank `dst`; returns None elsewhere.

    Uses a gloo (CPU) group: the data path needs no NCCL collective.
    """
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    lo, hi = shard_bounds(B, rank, world)
    assert local.shape[0] == hi - lo, (local.shape, lo, hi)
    per = -(-B // world)
    pad = np.zeros((per,) + local.shape[1:], dtype=local.dtype)
    pad[: hi - lo] = local
    t = torch.from_numpy(pad)
    if rank == dst:
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.gather(t, gather_list=bufs, dst=dst, group=group)
        out = np.empty((B,) + local.shape[1:], dtype=local.dtype)
        for r in range(world):
            l, h = shard_bounds(B, r, world)
            out[l:h] = bufs[r].numpy()[: h - l]
        return out
    dist.gather(t, gather_list=None, dst=dst, group=group)
    return None


def extract_multi_gpu(waves: np.ndarray, devices, extractor_kwargs: dict, **extract_kw):
    """Single-process variant: one host thread and one plan per device, outputs written
    into slices of shared host arrays (the "host gather").

    Every keyword of ``FeatureExtractor.extract_host`` is honoured (``pad_to``, ``pooled``, ``chroma``,
    ``fixed_frames``, ``sr_in``, ...): the shared arrays are sized from them, and each rank writes
    straight into its slice (``extract_host`` refuses a buffer of the wrong shape instead of
    replacing it, so nothing can land in a private array)."""
    from ._lib import lib
    from .core import FeatureExtractor, get_extractor

    waves = np.asarray(waves)
    B = waves.shape[0]
    world = len(devices)
    # one plan per DISTINCT device from the process-wide cache (plans keep their pinned-pipeline slots and lazily
    # built tables between calls); a device listed twice gets a private second plan, since a plan serves one
    # host thread at a time
    seen, exs, private = set(), [], []
    for d in devices:
        if d in seen:
            exs.append(FeatureExtractor(device=d, **extractor_kwargs))
            private.append(exs[-1])
        else:
            seen.add(d)
            exs.append(get_extractor(device=d, **extractor_kwargs))
    ex0 = exs[0]
    kw = dict(extract_kw)
    kw.pop("out", None)
    sr_in = int(kw.get("sr_in") or 0)
    if sr_in == ex0.params.sr:
        sr_in = 0
    n_res = int(lib.hlmc_resampled_length(waves.shape[1], sr_in, ex0.params.sr)) if sr_in else waves.shape[1]
    n_total = int(kw.get("pad_to") or n_res)
    T = ex0.num_frames(n_total)
    chroma = kw.get("chroma", False)
    has_mfcc = ex0.n_mfcc > 0
    pinned = kw.pop("pinned", True)

    def new(shape, dtype=np.float32):
        # the gather target: page-locked when torch can provide it (device-to-host copies into pageable memory
        # are staged by the driver and run at a fraction of the PCIe rate)
        if pinned and int(np.prod(shape)) > 0:
            try:
                import torch

                return torch.empty(shape, dtype=torch.float32 if dtype == np.float32 else torch.int32,
                                   pin_memory=True).numpy()
            except Exception:
                pass
        return np.empty(shape, dtype)

    out = {}
    if kw.get("logmel", True):
        out["logmel"] = new((B, ex0.n_mels, T))
    if kw.get("mfcc", True) and has_mfcc:
        out["mfcc"] = new((B, ex0.n_mfcc, T))
    if kw.get("stats", True):
        out["stats"] = new((B, 5, T))
    if kw.get("status", True):
        out["status"] = new((B,), np.int32)
    if kw.get("pooled", False):
        out["pooled"] = new((B, ex0.pooled_width(has_mfcc, bool(chroma))))
    if chroma and chroma != "pooled":
        out["chroma"] = new((B, 12, T))
        out["tuning"] = new((B,))
    if kw.get("fixed_frames"):
        out["fixed_logmel"] = new((B, ex0.n_mels, int(kw["fixed_frames"])))
    if kw.get("wave_out"):
        out["wave"] = new((B, n_total))
    errs = []

    def work(r):
        try:
            lo, hi = shard_bounds(B, r, world)
            if hi > lo:
                mine = {k: v[lo:hi] for k, v in out.items()}
                got = exs[r].extract_host(waves[lo:hi], out=mine, **kw)
                for k, v in mine.items():
                    assert got[k] is v, f"rank {r}: output {k!r} did not land in the shared array"
        except Exception as e:  # surfaced after join
            errs.append(e)

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    for ex in private:
        ex.close()
    if errs:
        raise errs[0]
    return out
