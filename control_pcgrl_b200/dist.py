"""Multi-GPU plumbing: env-index sharding (no collective on the step path) and the optional per-rollout
episode-stat reduction (the only cross-env reduction in the reference: RLlib merging
episode.custom_metrics from rollout workers, rl/callbacks.py:101-116).

One process per GPU (torchrun).  The step path never communicates: rank r owns envs
[r*N/W, (r+1)*N/W) and steps them with its own kernel launches.  `reduce_episode_stats` packs
[count, sum_k, sumsq_k] into ONE float64 vector (<= 1 KB) and issues one all-reduce(SUM) plus one
all-reduce(MAX) on [max_k, -min_k]: latency-bound on NVLink/NVSwitch, called once per rollout.
Works with the nccl backend on GPUs and with gloo on CPU (tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous env-index range of `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's env (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def bind_to_gpu_cpus(device_index: int) -> list:
    """One process per GPU: pin this process to the host cores NVML reports as local to the GPU (same NUMA node /
    PCIe root), intersected with the cores the container allows.  Pinned staging buffers allocated afterwards
    land on that node, so the host<->device copies of the e2e path do not cross sockets when 8 ranks run at
    once.  Returns the cores bound to ([] = left unchanged: NVML missing, or nothing in common)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        # torch's device index -> NVML handle through the PCI bus id (CUDA_VISIBLE_DEVICES may reorder)
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id if hasattr(
            torch.cuda.get_device_properties(device_index), "pci_bus_id") else None
        h = None
        if bus is not None:
            for i in range(pynvml.nvmlDeviceGetCount()):
                hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                if pynvml.nvmlDeviceGetPciInfo(hi).bus == bus:
                    h = hi
                    break
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cores = sorted(local & allowed)
        if cores and len(cores) < len(allowed):
            os.sched_setaffinity(0, cores)
            return cores
    except Exception:  # noqa: BLE001 -- affinity is an optimisation, never a requirement
        pass
    return []


def reduce_episode_stats(values: torch.Tensor, names=None, group=None):
    """values: [n_local_episodes, K] (any numeric dtype, any device) -> dict of global
    count / mean / std / min / max per column, identical on every rank."""
    v = values.to(torch.float64)
    if v.ndim == 1:
        v = v[:, None]
    k = v.shape[1]
    dev = v.device
    packed = torch.zeros(1 + 2 * k, dtype=torch.float64, device=dev)
    packed[0] = v.shape[0]
    if v.shape[0]:
        packed[1:1 + k] = v.sum(0)
        packed[1 + k:] = (v * v).sum(0)
        ext = torch.cat([v.max(0).values, -v.min(0).values])
    else:
        ext = torch.full((2 * k,), -float("inf"), dtype=torch.float64, device=dev)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
    cnt = packed[0].clamp(min=1)
    mean = packed[1:1 + k] / cnt
    var = (packed[1 + k:] / cnt - mean * mean).clamp(min=0)
    out = {"count": int(packed[0].item()), "mean": mean, "std": var.sqrt(), "max": ext[:k], "min": -ext[k:]}
    if names is not None:
        out["names"] = list(names)
    return out
