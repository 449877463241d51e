"""Instance sharding for multi-GPU runs: independent encrypted program instances are split into contiguous
ranges, one per rank (one process + one CudaCiphertextFactory per GPU, same key seed on every rank).
There is no data-path collective; torch.distributed is used only to agree on timings."""


def instance_range(n_instances, world_size, rank):
    """Contiguous [lo, hi) of the instances rank `rank` owns; sizes differ by at most one."""
    if not (0 <= rank < world_size) or n_instances < 0:
        raise ValueError("bad sharding arguments")
    base, extra = divmod(n_instances, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(values, dist=None, device="cpu"):
    """Element-wise maximum of a list of floats over all ranks (device timings are reported as max over ranks)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(values)
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def sum_over_ranks(values, dist=None, device="cpu"):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(values)
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]
