import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from bayesfast_b200 import _cabi
from oracle import bf_oracle
from _specs import synthetic_spec, to_device_spec
from test_gpu_sampler import cfg_from, device_draws
bf_oracle.build()
n, order, C, n_iter = 12, 'cubic-2', 37, 40
spec, cov = synthetic_spec(n, order, seed=90 + n)
spec['alpha'] = spec['alpha'] / 1.6
h = _cabi.Handle(0); h.set_model(to_device_spec(spec))
x0 = (np.linalg.cholesky(cov) @ np.random.default_rng(2).normal(size=(n, C))).T
cfg = cfg_from({'n_int_step': 12}, 20, 77, chain0=9)
res = {}
for mode in ('dmma', 'generic'):
    os.environ['BFB200_SAMPLER'] = mode
    h.sampler_init(cfg, x0, 0.5, np.ones(n), x0)
    res[mode] = h.sampler_run('HMC', n_iter)
    st = h.sampler_state()
    print(mode, h.sampler_last_path(), 'status', st['status'].max())
U, Z = device_draws(h, 77, st['n_draws'], 9)
ref = bf_oracle.OracleDensity(spec).run('HMC', dict(n_iter=n_iter, n_warmup=20, n_int_step=12), x0, 0.5, np.ones(n), draws_u=U, draws_z=Z)
for a, b in (('dmma', 'ref'), ('generic', 'ref'), ('dmma', 'generic')):
    A = res[a]['samples'] if a != 'ref' else ref['samples']; B = res[b]['samples'] if b != 'ref' else ref['samples']
    d = np.abs(A - B).max(axis=2)       # [C, iter]
    bad = np.argwhere(d > 1e-3)
    print(a, 'vs', b, 'max diff per iter', ['%.0e' % v for v in d.max(axis=0)], 'first bad (chain, iter)', bad[:3].tolist())
    if len(bad):
        c, i = bad[0]
        print('  chain', c, 'accepted', res['dmma']['tree_depth'][c, max(0,i-2):i+2], ref['tree_depth'][c, max(0,i-2):i+2], 'diverging', res['dmma']['diverging'][c, max(0,i-2):i+2], ref['diverging'][c, max(0,i-2):i+2],
              'dE', res['dmma']['energy_change'][c, max(0,i-2):i+2], ref['energy_change'][c, max(0,i-2):i+2], 'step', res['dmma']['step_size'][c, max(0,i-2):i+2], ref['step_size'][c, max(0,i-2):i+2])
