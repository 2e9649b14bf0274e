"""Dense mass matrix (metric='full', QuadMetricFullAdapt) on the generic warp-per-chain kernel against the diagonal default
on the same kernel and on the tensor-core kernel: d=26 cubic-2, full 1500-iteration NUTS runs (steady state, no outputs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N_ITER = int(sys.argv[2]) if len(sys.argv) > 2 else 1500          # a short run (e.g. 512 chains, 120 iterations) for ncu
ONLY_DENSE = '--dense-only' in sys.argv
n = 26
prob = synthetic.des_shaped(n, seed=1, n_chain=C)
sur = bfb.PolyModel('cubic-2', input_size=n, output_size=1)
sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
den = bfb.Density(sur)
h = den._sync(False)
cfg = bfb.NTrace(n_chain=C, n_iter=N_ITER, n_warmup=N_ITER // 3, x_0=prob['x_0'])._cfg_dict(1, 0)
for name, var0, dense, env in (('dense/generic', np.eye(n), True, None), ('diag/generic', np.ones(n), False, 'generic'),
                               ('diag/dmma', np.ones(n), False, None))[:1 if ONLY_DENSE else 3]:
    if env:
        os.environ['BFB200_SAMPLER'] = env
    else:
        os.environ.pop('BFB200_SAMPLER', None)
    h.sampler_init(cfg, prob['x_0'], 1. / n**0.25, var0, prob['x_0'], dense=dense)
    for rep in range(1 if ONLY_DENSE else 2):
        h.sampler_reset()
        r = h.sampler_run('NUTS', N_ITER, out_ptrs={})
        ms = h.last_kernel_ms()
    print('%-14s %s leaves %d ms %.1f rate %.3e mean tree size %.2f' % (name, h.sampler_last_path(), r['total_tree_size'], ms,
          r['total_tree_size'] / ms * 1e3, r['total_tree_size'] / (C * float(N_ITER))), flush=True)
