"""timing of sample() output modes (4096 chains, d=26): which part costs what"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic
n, C = 26, 4096
prob = synthetic.des_shaped(n, seed=1, n_chain=C)
sur = bfb.PolyModel('cubic-2', input_size=n, output_size=1)
sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
den = bfb.Density(sur)
kw = dict(n_chain=C, n_iter=1500, n_warmup=500, x_0=prob['x_0'], random_generator=3)
for name, o in (('all', dict()), ('post', dict(keep='post_warmup')), ('post_thin10', dict(keep='post_warmup', thin=10)),
                ('post_summ', dict(keep='post_warmup', summaries=True)), ('post_thin10_summ', dict(keep='post_warmup', thin=10, summaries=True)),
                ('post_thin10_samples_only', dict(keep='post_warmup', thin=10, fields=('samples', 'logp', 'tree_size')))):
    tt = None
    for i in range(3):
        del tt
        t0 = time.perf_counter()
        tt = bfb.sample(den, dict(kw), verbose=False, **o)
        dt = time.perf_counter() - t0
    print(name, 'wall ms %.1f kernel ms %.1f' % (dt * 1e3, tt.kernel_ms), 'bytes', sum(v.nbytes for k, v in tt.arrays.items() if 'original' not in k), flush=True)
