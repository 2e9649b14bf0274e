"""cubic-3 surrogate (d = 26: P = 3654) on the tensor-core path vs the generic kernels: evaluation, NUTS, HMC."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from bayesfast_b200 import _cabi
from _specs import synthetic_spec, to_device_spec
h = _cabi.Handle(0)
n, C = 26, 4096
spec, cov = synthetic_spec(n, 'cubic-3', seed=3, cubic_scale=0.02)
h.set_model(to_device_spec(spec))
peak = h.fp64_peak(0)
fl = 8 * n * n + 24 * n + 1.5 * n * (n - 1) * (n - 2)
Ce = 1 << 20
X = (torch.randn(Ce, n, dtype=torch.float64, device='cuda') @ torch.tensor(np.linalg.cholesky(cov).T, device='cuda')).contiguous()
lp = torch.empty(Ce, dtype=torch.float64, device='cuda'); g = torch.empty(Ce, n, dtype=torch.float64, device='cuda')
for mode in ('dmma', 'generic'):
    os.environ['BFB200_EVAL'] = mode
    ms = []
    for _ in range(4):
        h.logp_and_grad_batch_dev(X.data_ptr(), Ce, lp.data_ptr(), g.data_ptr()); ms.append(h.last_kernel_ms())
    print(json.dumps(dict(what='eval', mode=mode, ms=round(min(ms[1:]), 3), points_per_s=Ce / min(ms[1:]) * 1e3, frac_fp64=fl * Ce / min(ms[1:]) / 1e9 / peak)), flush=True)
x0 = (np.linalg.cholesky(cov) @ np.random.default_rng(0).normal(size=(n, C))).T
cfg = dict(n_warmup=150, max_treedepth=10, n_int_step=32, max_change=1000., adapt_step_size=1, target_accept=0.8, gamma=0.05, k=0.75,
           t0=10., adapt_metric=1, initial_weight=10., adapt_window=60, update_window=1, doubling=1, seed=1, chain0=0)
for sampler in ('NUTS', 'HMC'):
    for mode in ('dmma', 'generic'):
        os.environ['BFB200_SAMPLER'] = mode
        h.sampler_init(cfg, x0, 1. / n**0.25, np.ones(n), x0)
        h.sampler_run(sampler, 150, out_ptrs={})
        r = h.sampler_run(sampler, 100, out_ptrs={})
        ms = h.last_kernel_ms()
        print(json.dumps(dict(what=sampler, mode=h.sampler_last_path(), ms=round(ms, 2), leapfrogs_per_s=r['total_tree_size'] / ms * 1e3,
                              frac_fp64=fl * r['total_tree_size'] / ms / 1e9 / peak)), flush=True)
