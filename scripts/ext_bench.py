"""Extended density (decay + variable transform) on the tensor-core path vs the generic kernel: NUTS and HMC, 4096 chains."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from bayesfast_b200 import _cabi
from oracle import bf_oracle
from _specs import synthetic_spec, to_device_spec
bf_oracle.build()
n, C = 26, 4096
spec, cov = synthetic_spec(n, 'cubic-2', seed=1, decay=True, transform=True)
h = _cabi.Handle(0); h.set_model(to_device_spec(spec))
x0 = (np.linalg.cholesky(cov) @ np.random.default_rng(0).normal(size=(n, C))).T * 0.7
x0 = np.clip(x0, spec['transform_ranges'][:, 0] * 0.9, spec['transform_ranges'][:, 1] * 0.9)
x0 = np.array([bf_oracle.from_original(x, spec['transform_ranges'], spec['hard_bounds']) for x in x0])
cfg = dict(n_warmup=150, max_treedepth=10, n_int_step=32, max_change=1000., adapt_step_size=1, target_accept=0.8, gamma=0.05, k=0.75,
           t0=10., adapt_metric=1, initial_weight=10., adapt_window=60, update_window=1, doubling=1, seed=1, chain0=0)
for sampler in ('NUTS', 'HMC'):
    for mode in ('dmma', 'generic'):
        os.environ['BFB200_SAMPLER'] = mode
        h.sampler_init(cfg, x0, 1. / n**0.25, np.ones(n), x0)
        h.sampler_run(sampler, 150, fields=('tree_depth',))
        r = h.sampler_run(sampler, 150, fields=('tree_depth',))
        print(sampler, mode, h.sampler_last_path(), 'ms %.1f' % h.last_kernel_ms(), 'leapfrogs/s %.3e' % (r['total_tree_size'] / h.last_kernel_ms() * 1e3), flush=True)
