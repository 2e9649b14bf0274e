"""DES-Y1-shaped three-module density (n=27, m=457, shared 9-D mask, hard bounds, prior): evaluation and NUTS rates of the
feature-form tensor kernels against the dense-per-output tensor kernels and the generic ones."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic
p = synthetic.des_y1_like(457)
n, m = p['n'], p['m']
sur = bfb.PolyModel([bfb.PolyConfig('linear'), bfb.PolyConfig('quadratic', input_mask=p['nonlinear'])], input_size=n, output_size=m,
                    input_scales=p['ranges'])
pr = p['prior']
den = bfb.Density(sur, input_scales=p['ranges'], hard_bounds=True, likelihood=bfb.GaussianLikelihood(p['d'], np.ones(m), 0.),
                  prior=bfb.GaussianPrior(pr['idx'], pr['mu'], pr['sig']))
den.fit(p['x_fit'], p['y_fit'])
Pf = 1 + n + 45
C = 4096
x0 = p['x_0'][:C]
Xt = den.from_original(np.tile(x0, (16, 1)))
for name, env in (('feature', {}), ('dense', {'BFB200_LIK_DENSE': '1'}), ('generic', {'BFB200_EVAL': 'generic', 'BFB200_SAMPLER': 'generic'})):
    for k in list(os.environ):
        if k.startswith('BFB200_'):
            del os.environ[k]
    os.environ.update(env)
    den._dirty = True
    h = den._sync(False)
    peak = h.fp64_peak(0)
    for i in range(3):
        lp, g = den.logp_and_grad(Xt, original_space=False)
    ms = h.last_kernel_ms()
    tt = bfb.sample(den, dict(n_chain=C, n_iter=200 if name != 'generic' else 60, n_warmup=100 if name != 'generic' else 30, x_0=x0, random_generator=3), verbose=False,
                    fields=('tree_depth', 'samples', 'logp'))
    rate = tt.total_tree_size / tt.kernel_ms * 1e3
    print(json.dumps(dict(form=name, eval_path=h.eval_last_path(), sampler_path=h.sampler_last_path(), eval_points_per_s=Xt.shape[0] / ms * 1e3,
                          eval_frac_minimal=4 * m * Pf * Xt.shape[0] / ms / 1e9 / peak, nuts_leapfrogs_per_s=rate,
                          nuts_frac_minimal=4 * m * Pf * rate / 1e12 / peak, mean_depth=float(tt.arrays['tree_depth'].mean()),
                          logp_mean=float(lp.mean()))), flush=True)
