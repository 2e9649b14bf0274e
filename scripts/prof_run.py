"""Short sampler run for `ncu --set full`: d=26 cubic-2, warm-up launches then one launch to profile (use ncu -k regex: / -s).
usage: prof_run.py SAMPLER C n_iter [warm_iters]   (kernel family through BFB200_SAMPLER)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic

sampler = sys.argv[1]
C = int(sys.argv[2])
n_iter = int(sys.argv[3])
warm = int(sys.argv[4]) if len(sys.argv) > 4 else (500 if sampler == 'NUTS' else 100)
n = 26
prob = synthetic.des_shaped(n, seed=1, n_chain=C)
sur = bfb.PolyModel('cubic-2', input_size=n, output_size=1)
sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
den = bfb.Density(sur)
h = den._sync(False)
if sampler == 'HMC':
    cfg = bfb.HTrace(n_chain=C, n_iter=300, n_warmup=100, x_0=prob['x_0'], n_int_step=32)._cfg_dict(1, 0)
else:
    cfg = bfb.NTrace(n_chain=C, n_iter=1500, n_warmup=500, x_0=prob['x_0'])._cfg_dict(1, 0)
h.sampler_init(cfg, prob['x_0'], 1. / n**0.25, np.ones(n), prob['x_0'])
r = h.sampler_run(sampler, warm, fields=('tree_depth',))
print('warmup leaves', r['total_tree_size'], 'ms', h.last_kernel_ms(), h.sampler_last_path())
r = h.sampler_run(sampler, n_iter, fields=('tree_depth',))
print('profiled leaves', r['total_tree_size'], 'ms', h.last_kernel_ms(), 'rate', r['total_tree_size'] / h.last_kernel_ms() * 1e3)
