"""TNUTS on the device (csrc/bfb_sampler_tempered.cu): d = 26 cubic-2 target, quadratic surrogate of the same problem at three times
the temperature as the base density; leapfrogs/s of the warp-per-chain kernel (4 density evaluations per leapfrog: target and base
at the midpoint and at the end point) beside plain NUTS on the generic kernel for the same target."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic
n = 26
for C in (4096, 16384):
    prob = synthetic.des_shaped(n, seed=1, n_chain=C)
    sur = bfb.PolyModel('cubic-2', input_size=n, output_size=1)
    sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
    den = bfb.Density(sur)
    sb = bfb.PolyModel('quadratic', input_size=n, output_size=1)
    sb.fit(prob['x_fit'] * 1.5, prob['y_fit'] / 3., logp=prob['y_fit'][:, 0] / 3.)
    base = bfb.Density(sb)
    tr = bfb.TNTrace(base, 0., n_chain=C, n_iter=200, n_warmup=100, x_0=prob['x_0'], random_generator=3, u_0=np.zeros(C))
    tt = bfb.sample(den, tr, verbose=False, fields=('tree_size', 'tree_depth', 'u', 'weight'))
    rate = tt.total_tree_size / tt.kernel_ms * 1e3
    w = tt.arrays['weight'][:, 100:]
    print(json.dumps(dict(sampler='TNUTS', C=C, iters=200, kernel_ms=tt.kernel_ms, leapfrogs_per_s=rate, evaluations_per_s=4 * rate,
                          mean_tree_size=float(tt.arrays['tree_size'].mean()), mean_depth=float(tt.arrays['tree_depth'].mean()),
                          ess_weights=float(w.sum()**2 / (w**2).sum() / w.size), mean_u=float(tt.arrays['u'][:, 100:].mean()))), flush=True)
    os.environ['BFB200_SAMPLER'] = 'generic'
    t2 = bfb.sample(den, bfb.NTrace(n_chain=C, n_iter=200, n_warmup=100, x_0=prob['x_0'], random_generator=3), verbose=False,
                    fields=('tree_size',))
    del os.environ['BFB200_SAMPLER']
    print(json.dumps(dict(sampler='NUTS generic kernel', C=C, kernel_ms=t2.kernel_ms, leapfrogs_per_s=t2.total_tree_size / t2.kernel_ms * 1e3)), flush=True)
