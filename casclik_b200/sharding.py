"""Batch sharding across the GPUs of one box: one process per GPU, contiguous slices, no
collective on the data path (instances are independent — SURVEY.md §8e).  torch.distributed is
used only for the barrier / max-over-ranks timing and the optional final gather."""


def shard_range(n_total, rank, world):
    """Contiguous slice [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_columns(local, n_total, device=None):
    """All ranks contribute their (rows, n_local) slice; every rank gets (rows, n_total).
    Off the hot path: only for consumers that want all results in one place."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((local.shape[0], width), dtype=local.dtype, device=local.device)
    pad[:, :local.shape[1]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:, :hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=1)


def bind_to_device_numa_node(device_index):
    """Pin the calling process to the CPUs next to CUDA device `device_index` (NVML's ideal-CPU set
    for that GPU), so that page-locked host buffers allocated afterwards (first touch) and the
    zero-copy traffic of the host entry points stay on the GPU's own socket.  With one process per
    GPU this keeps 8 ranks from pulling each other's host buffers across the socket interconnect.
    Returns (previous affinity, new affinity) or None if NVML / the platform does not allow it."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            props = torch.cuda.get_device_properties(device_index)
            bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        before = sorted(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        after = sorted(os.sched_getaffinity(0))
        if not after:
            os.sched_setaffinity(0, before)
            return None
        return before, after
    except Exception:
        return None
