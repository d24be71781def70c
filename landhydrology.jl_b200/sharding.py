"""Column shards: the only multi-GPU structure on this path.

Soil columns are laterally independent (every reference RHS uses vertical operators only,
right_hand_side.jl:170-181, 337-365; BCs read the column's own cells 1 and n,
boundary_conditions.jl:182-185), so the column set is cut into contiguous ranges, one per
GPU/process, with no halo and no per-step exchange.  The single collective is the 2-double
water/energy budget all-reduce (``lh_soil_budgets_allreduce``: NCCL on the ctx stream).
``torch.distributed`` is used only as plumbing to hand rank 0's NCCL unique id to the others.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np


def shard_range(ncol: int, nranks: int, rank: int) -> Tuple[int, int]:
    """Contiguous range of rank ``rank``: [rank*ncol/nranks, (rank+1)*ncol/nranks)."""
    if not (0 <= rank < nranks):
        raise ValueError("rank out of range")
    if ncol < nranks:
        raise ValueError("fewer columns than ranks")
    return (rank * ncol) // nranks, ((rank + 1) * ncol) // nranks


@dataclass(frozen=True)
class ColumnShards:
    ncol: int
    nranks: int

    def ranges(self) -> List[Tuple[int, int]]:
        return [shard_range(self.ncol, self.nranks, r) for r in range(self.nranks)]

    def owner(self, col: int) -> int:
        for r, (lo, hi) in enumerate(self.ranges()):
            if lo <= col < hi:
                return r
        raise IndexError(col)

    def counts(self) -> List[int]:
        return [hi - lo for lo, hi in self.ranges()]


def exchange_unique_id(lib, dist) -> bytes:
    """Rank 0 creates the NCCL unique id; ``dist`` (torch.distributed, any backend) broadcasts it."""
    import torch

    rank = dist.get_rank()
    if rank == 0:
        uid = np.frombuffer(lib.comm_unique_id(), dtype=np.uint8).copy()
    else:
        uid = np.zeros(128, dtype=np.uint8)
    t = torch.from_numpy(uid)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def init_budget_comm(engine, dist) -> None:
    """Create the NCCL communicator used by ``lh_soil_budgets_allreduce`` on ``engine``'s ctx."""
    uid = exchange_unique_id(engine.lib, dist)
    engine.ctx.comm_init(dist.get_world_size(), dist.get_rank(), uid)


def allreduce_budgets_host(local: np.ndarray, dist) -> np.ndarray:
    """Host-side sum of per-rank budgets through ``dist`` (used by the gloo CPU tests to check
    the sharding arithmetic; the GPU path uses ``lh_soil_budgets_allreduce``)."""
    import torch

    t = torch.from_numpy(np.asarray(local, dtype=np.float64).copy())
    dist.all_reduce(t)
    return t.numpy()


def gpu_numa_cpus(device: int):
    """CPUs of the NUMA node the GPU ``device`` (CUDA ordinal of this process) hangs off, or ``None`` when the platform
    does not say (single-socket hosts, containers without sysfs).  Read from sysfs through the PCI bus id NVML reports."""
    try:
        import os

        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = device
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if device < len(ids) and ids[device].isdigit():
                index = int(ids[device])
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                   # NVML pads the domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return sorted(cpus) or None
    except Exception:
        return None


def bind_to_gpu_numa_node(device: int) -> bool:
    """Pin this process to the CPUs next to its GPU BEFORE it allocates pinned host buffers: page-locked memory is placed
    on the NUMA node of the allocating thread (first touch), and a buffer on the far socket halves the PCIe rate of an
    8-GPU box whose ranks all allocate from node 0 (e2e scaling).  Returns whether the affinity was changed."""
    import os

    cpus = gpu_numa_cpus(device)
    if not cpus:
        return False
    try:
        os.sched_setaffinity(0, cpus)
        return True
    except OSError:
        return False
