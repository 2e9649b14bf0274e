"""Short sampler run for `ncu --set full`: 4096 chains, d=26 cubic-2, one warm-up launch then one profiled launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 40
C = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
prob = synthetic.des_shaped(26, seed=1, n_chain=C)
sur = bfb.PolyModel('cubic-2', input_size=26, output_size=1)
sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
den = bfb.Density(sur)
h = den._sync(False)
cfg = bfb.NTrace(n_chain=C, n_iter=1500, n_warmup=500, x_0=prob['x_0'])._cfg_dict(1, 0)
h.sampler_init(cfg, prob['x_0'], 1. / 26**0.25, np.ones(26), prob['x_0'])
r = h.sampler_run('NUTS', 500, fields=('tree_depth',))           # adaptation phase (not profiled: use -k/-s)
print('warmup leaves', r['total_tree_size'], 'ms', h.last_kernel_ms())
r = h.sampler_run('NUTS', n_iter, fields=('tree_depth',))
print('profiled leaves', r['total_tree_size'], 'ms', h.last_kernel_ms(), 'rate', r['total_tree_size'] / h.last_kernel_ms() * 1e3)
