"""DES-Y1-shaped two-module pipeline: n = 26 parameters -> m block-quadratic surrogate outputs (masked
configs, examples/des-y1-w-cosmosis.ipynb cells 12-18) -> Gaussian likelihood with a dense inverse covariance.  Measures the
batched pipeline evaluation and a NUTS run on the generic kernels; the oracle port on the host cores runs beside it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from bayesfast_b200 import _cabi
from bayesfast_b200.density import whiten_spec, GaussianLikelihood
n = 26
m = int(sys.argv[1]) if len(sys.argv) > 1 else 457
C = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
N_IT = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 200      # a short run for ncu
rng = np.random.default_rng(0)
from bayesfast_b200 import synthetic
spec, lik = synthetic.des_pipeline(n, m, seed=0)     # 8 blocks of outputs, each a quadratic in 10 of the 26 inputs
ep = spec['epilogue']
h = _cabi.Handle(0)
h.set_model(whiten_spec(spec, lik))
X = rng.normal(size=(65536, n)) * 0.3
for rep in range(3):
    lp, g = h.logp_and_grad_batch(X)
    ms = h.last_kernel_ms()
print('pipeline eval: n=%d m=%d  %d points  %.2f ms  %.3e points/s' % (n, m, X.shape[0], ms, X.shape[0] / ms * 1e3), flush=True)
cfg = dict(n_warmup=N_IT // 2, max_treedepth=10, n_int_step=0, max_change=1000., adapt_step_size=1, target_accept=0.8, gamma=0.05,
           k=0.75, t0=10., adapt_metric=1, initial_weight=10., adapt_window=60, update_window=1, doubling=1, seed=1, chain0=0)
x0 = rng.normal(size=(C, n)) * 0.2
flops = m * (2. * n * n + 5. * n) + 9. * n          # algorithmic flops per evaluation / leapfrog (S_o x, f_o, gradient accumulation)
print('  = %.2f TFLOP/s algorithmic' % (X.shape[0] / ms * 1e3 * flops / 1e12), flush=True)
for fam in ('dmma', 'generic'):
    os.environ['BFB200_SAMPLER'] = fam
    for sampler, kw in (('NUTS', {}), ('HMC', {'n_int_step': 16})):
        if fam == 'generic' and sampler == 'HMC':
            continue
        h.sampler_init(dict(cfg, **kw), x0, 1. / n**0.25 if sampler == 'NUTS' else 0.2, np.ones(n), x0)
        r = h.sampler_run(sampler, N_IT, out_ptrs={})
        ms = h.last_kernel_ms()
        print('pipeline %s (%s kernel): %d chains x %d iterations  %.1f ms  %.3e leapfrogs/s = %.2f TFLOP/s  mean tree size %.2f' % (
            sampler, h.sampler_last_path(), C, N_IT, ms, r['total_tree_size'] / ms * 1e3, r['total_tree_size'] / ms * 1e3 * flops / 1e12,
            r['total_tree_size'] / (C * float(N_IT))), flush=True)
os.environ.pop('BFB200_SAMPLER')
if '--cpu' in sys.argv:
    from oracle import bf_oracle
    bf_oracle.build()
    od = bf_oracle.OracleDensity(spec)
    t0 = time.time()
    lpo, go = od.logp_and_grad_batch(X[:8192])
    dt = time.time() - t0
    print('oracle port, %d host threads: %.3e points/s; max rel diff logp %.1e grad %.1e' % (
        os.cpu_count(), 8192 / dt, np.max(np.abs(lp[:8192] - lpo) / np.abs(lpo)), np.max(np.abs(g[:8192] - go)) / np.max(np.abs(go))))
