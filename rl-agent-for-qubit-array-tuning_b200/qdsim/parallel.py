"""Multi-GPU plumbing of the env batch.

The path shards, it does not communicate: envs are independent, so rank r of W owns the contiguous env-id range
``shard_range(n_env, r, W)``, keeps those envs' model records resident on its own GPU and simulates their scans with no
collective on the step path.  The one exchange is off the step path: per-env episode statistics (return, length, final
distances -- the quantities src/qadapt/training/utils/metrics_logger.py:71-99 reports) are all-gathered to every rank
once per rollout, in env order, over NCCL (gloo in the CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def gpu_numa_cpus(device: int):
    """CPUs of the NUMA node the GPU hangs off, or None if the platform does not say.  PCI address from the CUDA device
    properties (so CUDA_VISIBLE_DEVICES remaps are honoured), NUMA node and CPU list from sysfs."""
    try:
        p = torch.cuda.get_device_properties(device)
        addr = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return cpus or None
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def bind_to_gpu_numa_node(device: int) -> bool:
    """Pin this process to the CPUs next to its GPU BEFORE it allocates pinned host buffers, so that those buffers (first
    touch) and the copy engine's traffic stay on the GPU's own socket.  With one process per GPU on a two-socket box,
    unbound ranks all land their staging buffers wherever the scheduler started them and the device-to-host image copies
    of half the GPUs cross the socket interconnect.  Returns False (and changes nothing) when the topology is unknown."""
    cpus = gpu_numa_cpus(device)
    if not cpus:
        return False
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return False
        os.sched_setaffinity(0, allowed)
        return True
    except (OSError, AttributeError):
        return False


def shard_range(n_env: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) env-id range of ``rank``; the first ``n_env % world`` ranks get one extra env."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(n_env, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_env: int, world: int) -> list[int]:
    return [shard_range(n_env, r, world)[1] - shard_range(n_env, r, world)[0] for r in range(world)]


def gather_episode_stats(local_stats: torch.Tensor, n_env: int, group=None) -> torch.Tensor:
    """All-gather ``local_stats`` ([n_local, k], this rank's envs in env order) into ``[n_env, k]`` on every rank.

    Shards may differ by one env, so the exchange pads to the largest shard (one collective, no size negotiation)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local_stats.shape[0] == n_env
        return local_stats
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_env, world)
    assert local_stats.shape[0] == sizes[rank], f"rank {rank} holds {local_stats.shape[0]} envs, expected {sizes[rank]}"
    k = local_stats.shape[1]
    pad = max(sizes)
    send = torch.zeros((pad, k), dtype=local_stats.dtype, device=local_stats.device)
    send[:sizes[rank]] = local_stats
    recv = torch.empty((world * pad, k), dtype=local_stats.dtype, device=local_stats.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, pad, k)
    return torch.cat([recv[r, :sizes[r]] for r in range(world)], dim=0)
