import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from bayesfast_b200 import _cabi
from _specs import synthetic_spec, to_device_spec
n, C, n_iter = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
spec, cov = synthetic_spec(n, 'cubic-2', seed=7 + n)
h = _cabi.Handle(0); h.set_model(to_device_spec(spec))
x0 = (np.linalg.cholesky(cov) @ np.random.default_rng(5).normal(size=(n, C))).T
cfg = dict(n_warmup=n_iter // 2, max_treedepth=10, n_int_step=0, max_change=1000., adapt_step_size=1, target_accept=0.8,
           gamma=0.05, k=0.75, t0=10., adapt_metric=1, initial_weight=10., adapt_window=60, update_window=1, doubling=1, seed=4242, chain0=0)
res = {}
for ck in (n_iter, 16, 7):
    os.environ['BFB200_CHUNK_ITERS'] = str(ck)
    h.sampler_init(cfg, x0, 1. / n**0.25, np.ones(n), x0)
    res[ck] = h.sampler_run('NUTS', n_iter)
    st = h.sampler_state()
    print('chunk', ck, 'leaves', res[ck]['total_tree_size'], 'status', st['status'][:4], 'draws', st['n_draws'][:4])
a = res[n_iter]
for ck in (16, 7):
    b = res[ck]
    bad = np.argwhere(a['tree_depth'] != b['tree_depth'])
    print('chunk', ck, 'first depth mismatch (chain, iter):', bad[:5].tolist(), 'max sample diff', np.nanmax(np.abs(a['samples'] - b['samples'])))
    d = np.abs(a['samples'] - b['samples']).max(axis=(0, 2))
    print('  per-iter max diff', ['%.1e' % v for v in d[:40]])
