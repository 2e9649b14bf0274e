"""torchrun check of the multi-GPU path: rows of the fit sharded over ranks + ONE NCCL all-reduce of the
packed partial-sum buffer must reproduce the single-GPU fit; chains sharded over ranks must reproduce the single-GPU
chains bit for bit (global chain ids -> same Philox streams).  Also times the sharded fit sweep (BASELINE config 5).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dist_check.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic
from bayesfast_b200.runtime import shard_bounds

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n = 12
prob = synthetic.des_shaped(n, seed=3, n_chain=64)
x, y = prob['x_fit'], prob['y_fit']
lo, hi = shard_bounds(x.shape[0], rank, world)
s_d = bfb.PolyModel('cubic-2', input_size=n, output_size=1, device=local)
s_d.fit(x[lo:hi], y[lo:hi], logp=y[lo:hi, 0], comm=True)          # sharded rows + NCCL all-reduce
s_1 = bfb.PolyModel('cubic-2', input_size=n, output_size=1, device=local)
s_1.fit(x, y, logp=y[:, 0])                                        # all rows on this GPU
errs = [float(np.max(np.abs(a._coef - b._coef)) / np.max(np.abs(b._coef))) for a, b in zip(s_d.configs, s_1.configs)]
ok_fit = max(errs) < 1e-10 and abs(s_d._alpha - s_1._alpha) < 1e-12 * s_1._alpha and np.allclose(s_d._hess, s_1._hess, rtol=1e-10) \
    and np.allclose(s_d._f_mu, s_1._f_mu, rtol=1e-9)
# every rank must hold bit-identical coefficients (they all solve the same reduced system)
mine = torch.tensor(np.concatenate([c._packed.ravel() for c in s_d.configs]), device='cuda')
allc = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(allc, mine)
ok_same = all(torch.equal(allc[0], t) for t in allc)
# sharded sampling
den = bfb.Density(s_1)
kw = dict(n_chain=64, n_iter=60, n_warmup=30, x_0=prob['x_0'], random_generator=11)
tt_d = bfb.sample(den, dict(kw), verbose=False, comm=True)
tt_1 = bfb.sample(den, dict(kw), verbose=False)
clo, chi = shard_bounds(64, rank, world)
ok_smp = np.array_equal(tt_d.samples, tt_1.samples[clo:chi]) and np.array_equal(tt_d.arrays['tree_depth'], tt_1.arrays['tree_depth'][clo:chi]) \
    and tt_d[0].chain_id == clo
# sharded TNUTS (tempered sampler, second density in a handle of its own): same chains as the single-GPU run, u_0 sharded like x_0
s_b = bfb.PolyModel('quadratic', input_size=n, output_size=1, device=local)
s_b.fit(x * 1.5, y / 3., logp=y[:, 0] / 3.)
base = bfb.Density(s_b)
u0 = np.linspace(-1., 1., 64)
tkw = dict(n_chain=64, n_iter=40, n_warmup=20, x_0=prob['x_0'], random_generator=11, u_0=u0)
tn_d = bfb.sample(den, bfb.TNTrace(base, 0.2, **tkw), verbose=False, comm=True)
tn_1 = bfb.sample(den, bfb.TNTrace(base, 0.2, **tkw), verbose=False)
ok_tmp = all(np.array_equal(tn_d.arrays[k], tn_1.arrays[k][clo:chi]) for k in ('samples', 'u', 'weight', 'tree_size')) \
    and tn_d[0].chain_id == clo and tn_d.sampler == 'TNUTS'
ok_smp = ok_smp and ok_tmp
# sharded fit sweep timing (d=32, N rows per rank resident on the device)
import ctypes as C
from bayesfast_b200 import _cabi, fit as bfit
res = []
sur = bfb.PolyModel('cubic-2', input_size=32, output_size=1, device=local)
h = sur._dev(); h.set_model(sur.to_spec(with_bound=False)); L = _cabi.lib()
g = torch.Generator(device='cuda').manual_seed(rank)
for Ntot in (100000, 1000000):
    Nl = Ntot // world
    xs = torch.randn(Nl, 32, dtype=torch.float64, device='cuda', generator=g)
    ys = (-(xs * xs).sum(1, keepdim=True) * 0.5).contiguous()
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    _cabi.check(L.bfb_fit_begin(h._h, None))
    _cabi.check(L.bfb_fit_accumulate(h._h, xs.data_ptr(), ys.data_ptr(), None, Nl, _cabi.BFB_DEVICE))
    t1 = time.perf_counter()
    bfit._allreduce_buffer(h, None)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    out = np.empty(sur.n_param); rr = C.c_double(0)
    _cabi.check(L.bfb_fit_solve(h._h, out.ctypes.data_as(_cabi._dp), C.byref(rr)))
    t3 = time.perf_counter()
    res.append(dict(N_total=Ntot, world=world, gram_s=t1 - t0, allreduce_s=t2 - t1, solve_s=t3 - t2,
                    buffer_mb=int(L.bfb_fit_buffer_size(h._h)) * 8 / 1e6))
if rank == 0:
    print(json.dumps(dict(world=world, fit_coef_err=errs, ok_fit=bool(ok_fit), ok_identical_across_ranks=bool(ok_same),
                          ok_sharded_sampling=bool(ok_smp), ok_sharded_tempered=bool(ok_tmp), fit_sweep=res)))
ok = torch.tensor([int(ok_fit and ok_same and ok_smp)], device='cuda')
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(ok.item()) == 1 else 1)
