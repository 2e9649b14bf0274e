"""64-D cubic-3 NUTS (BASELINE configs[3]) short run for profiling / timing"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic, _cabi
C3 = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
spec3, cov3 = synthetic.cubic3_stack(64, seed=3)
h3 = _cabi.Handle(0)
h3.set_model(spec3)
x03 = (np.linalg.cholesky(cov3) @ np.random.default_rng(0).normal(size=(64, C3))).T
cfg3 = bfb.NTrace(n_chain=C3, n_iter=300, n_warmup=100, x_0=x03, random_generator=1)._cfg_dict(1, 0)
h3.sampler_init(cfg3, x03, 1. / 64**0.25, np.ones(64), x03)
h3.sampler_run('NUTS', 60, out_ptrs={})
r3 = h3.sampler_run('NUTS', iters, fields=('tree_depth',))
ms3 = h3.last_kernel_ms()
print(json.dumps(dict(kernel=h3.sampler_last_path(), ms=ms3, leaves=r3['total_tree_size'], rate=r3['total_tree_size'] / ms3 * 1e3)))
X = np.ascontiguousarray(np.tile(x03, (64, 1)))
for i in range(3):
    h3.logp_and_grad_batch(X)
    print('eval points/s', X.shape[0] / h3.last_kernel_ms() * 1e3, h3.eval_last_path())
