"""Multi-GPU plumbing: one process per GPU, evidence sets sharded across ranks, no collective on the
inference path; the only collective is the final posterior gather (NCCL all-gather over NVLink on GPUs,
gloo in the CPU tests). Evidence sets never interact, so plans and tables are replicated."""
from __future__ import annotations

from typing import Tuple


def shard_range(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `n` evidence sets for `rank`; sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows(x, world_size: int, rank: int):
    lo, hi = shard_range(x.shape[0], world_size, rank)
    return x[lo:hi]


def gather_posteriors(local, total_rows: int = None, group=None):
    """All-gather the per-rank posterior rows [B_r, out_elems] into [sum B_r, out_elems] on every rank,
    in rank order (ragged shards are padded to the largest shard for the collective and trimmed after)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rows = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    all_rows = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(all_rows, rows, group=group)
    counts = [int(r.item()) for r in all_rows]
    width = local.shape[1]
    biggest = max(counts)
    padded = local
    if local.shape[0] != biggest:
        padded = torch.zeros((biggest, width), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    full = torch.empty((world * biggest, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, padded.contiguous(), group=group)
    if all(c == biggest for c in counts):
        out = full
    else:
        out = torch.cat([full[r * biggest : r * biggest + counts[r]] for r in range(world)], dim=0)
    if total_rows is not None and out.shape[0] != total_rows:
        raise RuntimeError("gathered row count does not match the batch")
    return out


def gather_posteriors_to_root(local, dst: int = 0, group=None):
    """Gather the per-rank posterior rows [B_r, out_elems] on rank `dst` only (rank order; equal shards): every other
    rank sends its shard once over NVLink and receives nothing — the all-gather above moves world x more bytes to give
    every rank a copy nobody asked for. Returns [sum B_r, out_elems] on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if rank == dst:
        full = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        dist.gather(local.contiguous(), gather_list=list(full.unbind(0)), dst=dst, group=group)
        return full.view(world * local.shape[0], *local.shape[1:])
    dist.gather(local.contiguous(), gather_list=None, dst=dst, group=group)
    return None


def bind_process_to_gpu_numa(device_index: int):
    """Pin this process to the CPUs local to its GPU (PCIe root / NUMA node) so that pinned host buffers are
    first-touched on the right node; with 8 ranks streaming posteriors back to the host this decides whether the
    D2H copies fight over one socket. Returns the CPU list used, or None when the topology cannot be read."""
    import os

    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        handle = None
        for i in range(pynvml.nvmlDeviceGetCount()):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            u = pynvml.nvmlDeviceGetUUID(h)
            u = u.decode() if isinstance(u, bytes) else u
            if uuid in u or u.replace("GPU-", "") == uuid:
                handle = h
                break
        if handle is None:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        bus = pynvml.nvmlDeviceGetPciInfo(handle).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml pads the domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        path = f"/sys/bus/pci/devices/{bus}/local_cpulist"
        with open(path) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return text
    except Exception:
        return None
