"""Per-leapfrog latency of one group of 8 chains (HMC, d=26 cubic-2): one group per SM so nothing overlaps."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic
n = 26
C = 8 * 148
prob = synthetic.des_shaped(n, seed=1, n_chain=C)
sur = bfb.PolyModel('cubic-2', input_size=n, output_size=1)
sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
den = bfb.Density(sur)
h = den._sync(False)
for fam, env in (('team', {'BFB200_TEAMS_PER_SM': '1'}), ('team', {'BFB200_TEAMS_PER_SM': '4'}), ('dmma', {})):
    for k in list(os.environ):
        if k.startswith('BFB200_'):
            del os.environ[k]
    os.environ['BFB200_SAMPLER'] = fam
    os.environ['BFB200_CHUNK_ITERS'] = '100000'
    os.environ.update(env)
    cfg = bfb.HTrace(n_chain=C, n_iter=300, n_warmup=100, x_0=prob['x_0'], n_int_step=32)._cfg_dict(1, 0)
    h.sampler_init(cfg, prob['x_0'], 1. / n**0.25, np.ones(n), prob['x_0'])
    for k in (100, 200):
        r = h.sampler_run('HMC', k, fields=('tree_depth',))
        ms = h.last_kernel_ms()
        print(json.dumps(dict(family=fam, env=env, iters=k, ms=round(ms, 3), cycles_per_leapfrog=round(ms * 1e-3 * 1.965e9 / (k * 32), 1))), flush=True)
