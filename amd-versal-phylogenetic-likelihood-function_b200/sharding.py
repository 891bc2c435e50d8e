"""Multi-GPU plumbing for the PLF path: site-range sharding and the one optional collective.

Sites are independent, so the path shards with NO steady-state inter-GPU traffic: rank r of W
owns a contiguous site range chosen by the reference's instance-split rule
(app/src/include.h:181-192: ceil(n/W) each, the last takes the remainder).  The only exchange is
the final sum of the per-rank scaler increments (host_mem.cpp:384-388 is the single-device
version), an all-reduce of one int64 over NCCL (NVLink/NVSwitch) -- or gloo in the CPU tests.
"""
from __future__ import annotations

import math


def shard_for_rank(n_sites: int, rank: int, world: int):
    """(first_site, count) of `rank`; count may be 0 when n_sites < world."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    per = math.ceil(n_sites / world)
    lo = min(rank * per, n_sites)
    return lo, max(0, min(per, n_sites - lo))


def all_shards(n_sites: int, world: int):
    return [shard_for_rank(n_sites, r, world) for r in range(world)]


def reduce_scaler_increment(local_increment: int, device=None, group=None) -> int:
    """Sum of the per-rank scaler increments.  One int64 all-reduce; identity when
    torch.distributed is not initialised (single process)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return int(local_increment)
    t = torch.tensor([int(local_increment)], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def reduce_log_likelihood(local_lnl: float, device=None, group=None) -> float:
    """Sum of the per-rank log-likelihoods (one float64 all-reduce): sites are independent, so the
    log-likelihood of an alignment is the sum over the site ranges of the ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(local_lnl)
    t = torch.tensor([float(local_lnl)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Max of a per-rank scalar (device timings are reported as the max over ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def bind_host_to_device(device_index: int) -> dict:
    """Pin the calling process to the CPUs NVML reports as local to GPU `device_index`, so that pinned
    host buffers allocated afterwards land on the GPU's own NUMA node (first touch) and the copy
    engines do not cross the socket interconnect.  Pure plumbing for the host<->device legs of a
    multi-GPU box; a no-op (returns {"bound": False, ...}) when NVML, the affinity mask or
    sched_setaffinity are unavailable, or when the mask would be empty inside this cgroup."""
    import os
    info = {"bound": False, "device": device_index}
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if visible:
            entry = visible.split(",")[device_index].strip()
            if entry.isdigit():
                idx = int(entry)
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64 + 1)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        info["gpu_local_cpus"] = len(cpus)
        info["allowed_cpus"] = len(allowed)
        if target and target != allowed:
            os.sched_setaffinity(0, target)
            info["bound"] = True
        info["cpus"] = len(target or allowed)
    except Exception as e:  # no NVML / not permitted: leave the process where it is
        info["error"] = f"{type(e).__name__}: {e}"[:120]
    return info
