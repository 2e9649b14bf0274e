"""Run-to-run variance of the full 4096-chain NUTS launch inside one process (same module addresses, new launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic
C = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
prob = synthetic.des_shaped(26, seed=1, n_chain=C)
sur = bfb.PolyModel('cubic-2', input_size=26, output_size=1)
sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
den = bfb.Density(sur)
h = den._sync(False)
cfg = bfb.NTrace(n_chain=C, n_iter=1500, n_warmup=500, x_0=prob['x_0'])._cfg_dict(1, 0)
h.sampler_init(cfg, prob['x_0'], 1. / 26**0.25, np.ones(26), prob['x_0'])
ms = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    h.sampler_reset()
    r = h.sampler_run('NUTS', 1500, out_ptrs={})
    ms.append(round(h.last_kernel_ms(), 1))
print('in-process launches ms:', ms)
