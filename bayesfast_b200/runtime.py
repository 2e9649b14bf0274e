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


def _coll_device(group, device):
    """tensors of a collective live on the GPU for nccl, on the host for gloo (CPU tests, single-GPU multi-process runs)"""
    import torch.distributed as dist
    return 'cuda:{}'.format(int(device)) if dist.get_backend(group) == 'nccl' else 'cpu'


def allreduce_values(values, op, group, device, dtype='float64'):
    """all-reduce of a few host numbers (op: 'sum' | 'max'); returns a numpy array.  Used for row counts and radii of a fit
    whose rows are sharded over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    t = torch.tensor(np.atleast_1d(values), dtype=getattr(torch, dtype), device=_coll_device(group, device))
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 'max' else dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def allgather_vector(vec, group, device):
    """every rank's copy of a small float64 vector, stacked [world, len]"""
    import numpy as np
    import torch
    import torch.distributed as dist
    t = torch.tensor(np.asarray(vec, dtype=np.float64), device=_coll_device(group, device))
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return torch.stack(out).cpu().numpy()


def broadcast_vector(vec, src, group, device):
    import numpy as np
    import torch
    import torch.distributed as dist
    t = torch.tensor(np.asarray(vec, dtype=np.float64), device=_coll_device(group, device))
    dist.broadcast(t, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    return t.cpu().numpy()


def pick_center(best, x_best, group, device):
    """PolyModel._set_bound's center_max point (poly.py:277-287) over sharded rows: every rank contributes its best local
    (logp, x) -- or -inf when it has no valid candidate -- and all ranks pick the same winner; None when no rank has one."""
    import numpy as np
    allc = allgather_vector(np.concatenate(([best], x_best)), group, device)
    i = int(np.argmax(allc[:, 0]))
    return None if not np.isfinite(allc[i, 0]) else allc[i, 1:]


def shard_bounds(total, rank, world):
    """contiguous shard [lo, hi) of `total` items for `rank` of `world` (first `total % world` ranks get one more)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bind_to_gpu_numa(device):
    """Pin this process (and hence its first-touch / page-locked host buffers) to the CPUs of the NUMA node the GPU hangs off.
    With one process per GPU the device-to-host output copies of all ranks otherwise land on whatever node the scheduler
    picked (measured on an 8 x B200 box: end-to-end 260 ms per step against 161 ms for a single rank).  Best effort:
    returns the CPU list used, or None when the topology cannot be read."""
    import subprocess
    try:
        bus = subprocess.run(['nvidia-smi', '-i', str(int(device)), '--query-gpu=pci.bus_id', '--format=csv,noheader'],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if not bus:
            return None
        if bus.count(':') == 2 and len(bus.split(':')[0]) == 8:
            bus = bus[4:]                                   # 00000000:1B:00.0 -> 0000:1b:00.0
        base = '/sys/bus/pci/devices/' + bus
        cpulist = open(base + '/local_cpulist').read().strip()
        cpus = set()
        for part in cpulist.split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = set(os.sched_getaffinity(0))
        cpus &= allowed
        if not cpus or cpus == allowed:
            return sorted(cpus) if cpus else None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None
