"""End-to-end time of sample(keep='post_warmup') against the number of output chunks (BFB200_E2E_CHUNKS), bench configuration."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic

C, n = 4096, 26
prob = synthetic.des_shaped(n, seed=1, n_chain=C)
sur = bfb.PolyModel('cubic-2', input_size=n, output_size=1)
sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
den = bfb.Density(sur)
kw = dict(n_chain=C, n_iter=1500, n_warmup=500, x_0=prob['x_0'], random_generator=7)
for chunks in sys.argv[1:] or ['6', '10', '16']:
    os.environ['BFB200_E2E_CHUNKS'] = chunks
    for rep in range(4):
        t0 = time.perf_counter()
        tt = bfb.sample(den, dict(kw), verbose=False, keep='post_warmup')
        t1 = time.perf_counter()
        if rep:
            print('chunks %s: sample() total %.1f ms, kernel-span %.1f ms' % (chunks, (t1 - t0) * 1e3, tt.kernel_ms), flush=True)
        del tt
