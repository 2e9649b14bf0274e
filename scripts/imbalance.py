"""How unevenly the leapfrogs are spread over the 8-chain groups (the unit a warp owns) and over single chains."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prob = synthetic.des_shaped(26, seed=1, n_chain=C)
sur = bfb.PolyModel('cubic-2', input_size=26, output_size=1)
sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
den = bfb.Density(sur)
h = den._sync(False)
cfg = bfb.NTrace(n_chain=C, n_iter=1500, n_warmup=500, x_0=prob['x_0'])._cfg_dict(1, 0)
h.sampler_init(cfg, prob['x_0'], 1. / 26**0.25, np.ones(26), prob['x_0'])
r = h.sampler_run('NUTS', 1500, fields=('tree_size',))
ts = r['tree_size'].astype(np.int64)            # [C, 1500]
print('kernel ms', h.last_kernel_ms(), 'leaves', ts.sum(), 'rate %.3e' % (ts.sum() / h.last_kernel_ms() * 1e3))
for name, sl in (('warmup 0-500', slice(0, 500)), ('sampling 500-1500', slice(500, 1500)), ('all', slice(0, 1500))):
    t = ts[:, sl]
    per_chain = t.sum(axis=1)
    chunk = 250
    nchunk = t.shape[1] // chunk
    g = t.reshape(C // 8, 8, nchunk, chunk).sum(axis=3)          # leaves per (group, chain, chunk)
    rounds_g = g.max(axis=1).sum(axis=1)                          # rounds a warp needs for its group: sum over chunks of the slowest chain
    print(name, 'per chain: mean %.0f max %.0f (x%.2f)' % (per_chain.mean(), per_chain.max(), per_chain.max() / per_chain.mean()),
          '| rounds per group: mean %.0f max %.0f (x%.2f)' % (rounds_g.mean(), rounds_g.max(), rounds_g.max() / rounds_g.mean()),
          '| row utilisation %.3f' % (g.sum() / (8 * rounds_g.sum())))
    it_max = t.reshape(C // 8, 8, -1).max(axis=1).sum()           # if chains of a group were in lock step per iteration
    print('   lock-step-per-iteration utilisation would be %.3f' % (t.sum() / (8 * it_max)))
first = ts[:, :50]
print('mean tree size, iterations 0-9:', first[:, :10].mean(axis=0).round(1).tolist())
print('max tree size,  iterations 0-9:', first[:, :10].max(axis=0).tolist())
