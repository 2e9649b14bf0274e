"""Process-level plumbing: which GPU this process drives, torch.distributed helpers (one process per GPU)."""
import os


def default_device():
    """LOCAL_RANK under torchrun, else BFB200_DEVICE, else 0."""
    for k in ('LOCAL_RANK', 'BFB200_DEVICE'):
        v = os.environ.get(k)
        if v is not None:
            try:
                return int(v)
            except ValueError:
                pass
    return 0


def dist_info(comm=None):
    """(rank, world_size, group) of the torch.distributed group in use, or (0, 1, None)."""
    if comm is None or comm is False:
        return 0, 1, None
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1, None
    group = None if comm is True else comm
    return dist.get_rank(group), dist.get_world_size(group), group


def shard_bounds(total, rank, world):
    """contiguous shard [lo, hi) of `total` items for `rank` of `world` (first `total % world` ranks get one more)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
