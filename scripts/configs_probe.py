"""One number for each BASELINE.json configuration that is not the bench line (steady state, outputs not written)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from bayesfast_b200 import _cabi
from _specs import synthetic_spec, to_device_spec
h = _cabi.Handle(0)
base = dict(n_warmup=100, max_treedepth=10, n_int_step=0, max_change=1000., adapt_step_size=1, target_accept=0.8, gamma=0.05, k=0.75,
            t0=10., adapt_metric=1, initial_weight=10., adapt_window=60, update_window=1, doubling=1, seed=1, chain0=0)
for name, n, order, C, iters in (('configs[0] 2-D quadratic, 4 chains', 2, 'quadratic', 4, 500),
                                 ('configs[1] 16-D cubic-2, 4096 chains', 16, 'cubic-2', 4096, 200),
                                 ('configs[3] 64-D cubic-3, 1024 chains', 64, 'cubic-3', 1024, 40)):
    spec, cov = synthetic_spec(n, order, seed=3, bound=(n != 64))
    h.set_model(to_device_spec(spec))
    x0 = (np.linalg.cholesky(cov) @ np.random.default_rng(0).normal(size=(n, C))).T
    h.sampler_init(base, x0, 1. / n**0.25, np.ones(n), x0)
    h.sampler_run('NUTS', 100, out_ptrs={})
    r = h.sampler_run('NUTS', iters, fields=('tree_depth',))
    r = h.sampler_run('NUTS', iters, fields=('tree_depth',))
    ms = h.last_kernel_ms()
    print(json.dumps(dict(config=name, kernel=h.sampler_last_path(), ms=round(ms, 2), leapfrogs_per_s=r['total_tree_size'] / ms * 1e3,
                          mean_depth=float(r['tree_depth'].mean()), depth_hist=np.bincount(r['tree_depth'].ravel(), minlength=11).tolist())), flush=True)
