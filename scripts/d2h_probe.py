"""What bounds the end-to-end run on 8 GPUs: concurrent device-to-host bandwidth of one host for the copy patterns of
bfb_sampler_run_ex -- (a) one contiguous copy, (b) the samples' strided chunk copies (4096 rows of 62 x 208 B), (c) the statistics'
strided chunk copies (4096 rows of 62 x 8 B, ten fields) -- with only rank 0 copying and with all ranks copying at once.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 scripts/d2h_probe.py
"""
import ctypes as C
import glob
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), '..', 'nvidia', 'cuda_runtime', 'lib', 'libcudart.so*')) + \
    glob.glob('/usr/local/cuda/lib64/libcudart.so*')
rt = C.CDLL(cands[0])
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
D2H = 2
Cc, R, n, nch = 4096, 1000, 26, 16
K = (R + nch - 1) // nch
dev = torch.empty(Cc * R * n, dtype=torch.float64, device='cuda')
host = torch.empty(Cc * R * n, dtype=torch.float64).pin_memory()
dstat = torch.empty(10 * Cc * R, dtype=torch.float64, device='cuda')
hstat = torch.empty(10 * Cc * R, dtype=torch.float64).pin_memory()
st = torch.cuda.Stream()
sp = C.c_void_p(st.cuda_stream)


def contiguous():
    rt.cudaMemcpyAsync(host.data_ptr(), dev.data_ptr(), dev.numel() * 8, D2H, sp)
    return dev.numel() * 8


def samples_chunks():
    fb = n * 8
    for k in range(nch):
        its = min(K, R - k * K)
        rt.cudaMemcpy2DAsync(host.data_ptr() + fb * k * K, fb * R, dev.data_ptr() + fb * k * K, fb * R, fb * its, Cc, D2H, sp)
    return dev.numel() * 8


def stats_chunks():
    fb = 8
    for k in range(nch):
        its = min(K, R - k * K)
        for f in range(10):
            off = f * Cc * R * 8
            rt.cudaMemcpy2DAsync(hstat.data_ptr() + off + fb * k * K, fb * R, dstat.data_ptr() + off + fb * k * K, fb * R, fb * its, Cc, D2H, sp)
    return dstat.numel() * 8


def stats_whole():
    rt.cudaMemcpyAsync(hstat.data_ptr(), dstat.data_ptr(), dstat.numel() * 8, D2H, sp)
    return dstat.numel() * 8


def timed(fn, active):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = 0
    if active:
        for _ in range(3):
            b += fn()
        st.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt if active else 0., float(b)], device='cuda', dtype=torch.float64)
    if world > 1:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        return float(sm[1]) / float(mx[0]) / 1e9
    return b / dt / 1e9


res = {}
for name, fn in (('contiguous', contiguous), ('samples_chunks', samples_chunks), ('stats_chunks', stats_chunks), ('stats_whole', stats_whole)):
    timed(fn, True)                                           # warm
    res[name] = dict(rank0_alone_gbs=timed(fn, rank == 0), all_ranks_aggregate_gbs=timed(fn, True))
if rank == 0:
    print(json.dumps(dict(world=world, cpus=os.cpu_count(), **res)))
if world > 1:
    dist.destroy_process_group()
