"""First-contact probe: FP64 peaks (DFMA / DMMA / both) and a quick NUTS throughput number."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from bayesfast_b200 import _cabi
from _specs import synthetic_spec, to_device_spec

h = _cabi.Handle(0)
out = {}
for kind, name in ((0, 'dfma'), (1, 'dmma'), (2, 'both')):
    out[name + '_tflops'] = [round(h.fp64_peak(kind), 2) for _ in range(3)]
print(json.dumps(out))
for n, order, C in ((26, 'cubic-2', 4096), (26, 'cubic-2', 16384), (16, 'cubic-2', 4096)):
    spec, cov = synthetic_spec(n, order, seed=1)
    h.set_model(to_device_spec(spec))
    x0 = (np.linalg.cholesky(cov) @ np.random.default_rng(0).normal(size=(n, C))).T
    cfg = dict(n_warmup=100, max_treedepth=10, n_int_step=0, max_change=1000., adapt_step_size=1, target_accept=0.8,
               gamma=0.05, k=0.75, t0=10., adapt_metric=1, initial_weight=10., adapt_window=60, update_window=1,
               doubling=1, seed=1, chain0=0)
    h.sampler_init(cfg, x0, 1. / n**0.25, np.ones(n), x0)
    r = h.sampler_run('NUTS', 100, fields=('tree_depth',))
    ms = h.last_kernel_ms()
    r2 = h.sampler_run('NUTS', 100, fields=('tree_depth',))
    ms2 = h.last_kernel_ms()
    print(json.dumps(dict(n=n, C=C, warm_leaf=r['total_tree_size'], warm_ms=ms, warm_rate=r['total_tree_size'] / ms * 1e3,
                          post_leaf=r2['total_tree_size'], post_ms=ms2, post_rate=r2['total_tree_size'] / ms2 * 1e3,
                          mean_depth=float(r2['tree_depth'].mean()))))
