"""Where the end-to-end time of bayesfast_b200.sample() goes (host side), bench configuration."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic, _cabi

C, n = 4096, 26
prob = synthetic.des_shaped(n, seed=1, n_chain=C)
sur = bfb.PolyModel('cubic-2', input_size=n, output_size=1)
sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
den = bfb.Density(sur)
kw = dict(n_chain=C, n_iter=1500, n_warmup=500, x_0=prob['x_0'], random_generator=7)
for rep in range(3):
    t0 = time.perf_counter()
    tt = bfb.sample(den, dict(kw), verbose=False)
    t1 = time.perf_counter()
    print('sample() total %.1f ms, kernel-span %.1f ms, leaves %d' % ((t1 - t0) * 1e3, tt.kernel_ms, tt.total_tree_size))
    del tt
h = den._sync(False)
trace = bfb.NTrace(**kw)
cfg = trace._cfg_dict(7, 0)
x0 = np.ascontiguousarray(prob['x_0'])
for rep in range(2):
    t0 = time.perf_counter(); h.sampler_init(cfg, x0, 1. / n**0.25, np.ones(n), x0); t1 = time.perf_counter()
    res = h.sampler_run('NUTS', 1500); t2 = time.perf_counter()
    st = h.sampler_state(); t3 = time.perf_counter()
    print('init %.1f ms | run %.1f ms (kernel span %.1f) | state %.1f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, h.last_kernel_ms(), (t3 - t2) * 1e3))
    del res
# device-resident outputs, single launch
import torch
S = C * 1500
outs = dict(samples=torch.empty(S * n, dtype=torch.float64, device='cuda'))
for k in _cabi.FLOAT_STATS: outs[k] = torch.empty(S, dtype=torch.float64, device='cuda')
for k in _cabi.INT_STATS: outs[k] = torch.empty(S, dtype=torch.int32, device='cuda')
h.sampler_init(cfg, x0, 1. / n**0.25, np.ones(n), x0)
t0 = time.perf_counter(); r = h.sampler_run('NUTS', 1500, out_ptrs={k: v.data_ptr() for k, v in outs.items()}); t1 = time.perf_counter()
print('device-resident single launch: %.1f ms (kernel %.1f)' % ((t1 - t0) * 1e3, h.last_kernel_ms()))
