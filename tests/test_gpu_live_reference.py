"""The drop-in boundary driven by the REAL reference, live: a bayesfast.Density (vendored, unmodified, by __graft_entry__.build()
into baseline/_ref, which travels to the GPU box with the snapshot) is fitted by the reference itself, handed to
bayesfast_b200.Density.from_reference / bayesfast_b200.sample, and the device's logp / gradient are compared with the reference's own
Density.logp_and_grad at the same points (core/density.py:724-754), the device's fit with the reference's PolyModel.fit
(modules/poly.py:505-589).  Runs in a subprocess: the import shim of the reference (numpy-2 aliases, stub matplotlib) must not leak
into the test process.  Skipped when baseline/_ref was not built."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref', 'bayesfast')

SCRIPT = r'''
import json, sys, os, warnings
import numpy as np
ROOT = sys.argv[1]
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'baseline'))
import ref_runner
bf = ref_runner.import_reference()
from bayesfast.modules.poly import PolyModel
import bayesfast_b200 as bfb
warnings.simplefilter('ignore')
out = {}
rng = np.random.default_rng(31)
for name, n, order, transform, decay in (('c2_n6_transform_decay', 6, 'cubic-2', True, True), ('c2_n26_bound', 26, 'cubic-2', False, False),
                                         ('c3_n5_decay', 5, 'cubic-3', False, True)):
    A = rng.normal(size=(n, n))
    cov = A @ A.T / n + np.eye(n)
    P = np.linalg.inv(cov)
    target = lambda x: np.atleast_1d(-0.5 * x @ P @ x - 0.02 * np.sum(x**3 * np.exp(-0.1 * x**2)))
    mod = bf.Module(fun=target, input_vars='x', output_vars='logp')
    sur = PolyModel(order, input_size=n, output_size=1, input_vars='x', output_vars='logp')
    kw = {}
    if transform:
        ranges = np.stack((-12. - rng.uniform(size=n), 12. + rng.uniform(size=n)), axis=1)
        hb = np.zeros((n, 2), np.uint8); hb[0] = (1, 1); hb[2] = (1, 0); hb[3] = (0, 1)
        kw.update(input_scales=ranges, hard_bounds=hb)
    den = bf.Density(density_name='logp', module_list=[mod], surrogate_list=[sur], input_vars='x',
                     decay_options={'use_decay': decay}, **kw)
    L = np.linalg.cholesky(cov)
    xf = (L @ rng.normal(size=(n, 4 * sur.n_param))).T
    if transform:
        xf = np.clip(xf, -11., 11.)
    den.fit([den.fun(x, original_space=True, use_surrogate=False) for x in xf])          # the reference's own fit
    den.use_surrogate = True
    # points inside and far outside the radial bound, in the transformed space
    Xo = np.concatenate(((L @ rng.normal(size=(n, 24))).T * 0.8, (L @ rng.normal(size=(n, 8))).T * 4.))
    if transform:
        Xo = np.clip(Xo, -11.5, 11.5)
    Xt = np.array([den.from_original(x) for x in Xo])
    lp_ref = np.array([float(den.logp_and_grad(x, original_space=False)[0]) for x in Xt])
    g_ref = np.array([np.asarray(den.logp_and_grad(x, original_space=False)[1]) for x in Xt])
    dd = bfb.Density.from_reference(den)                                                     # the drop-in boundary
    lp, g = dd.logp_and_grad(Xt, original_space=False)
    e_lp = float(np.max(np.abs(lp - lp_ref) / np.maximum(np.abs(lp_ref), 1e-3 * np.abs(lp_ref).max())))
    e_g = float(np.max(np.abs(g - g_ref) / np.maximum(np.abs(g_ref), 1e-3 * np.abs(g_ref).max())))
    # the device's own fit of the same rows against the reference's coefficients
    y = np.array([float(target(x)[0]) for x in xf])[:, None]
    s2 = bfb.PolyModel(order, input_size=n, output_size=1)
    s2.fit(xf if not transform else xf, y, logp=y[:, 0])
    e_fit = 0.
    for c_dev, c_ref in zip(s2.configs, sur._configs):
        a, b = np.asarray(c_dev._coef, dtype=float), np.asarray(c_ref._coef, dtype=float)
        mask = np.isfinite(b)
        if c_ref.order == 'quadratic':
            mask &= np.triu(np.ones(b.shape[1:], bool))[None]
        if c_ref.order == 'cubic-3':
            i = np.arange(b.shape[1]); mask &= ((i[:, None, None] < i[None, :, None]) & (i[None, :, None] < i[None, None, :]))[None]
        e_fit = max(e_fit, float(np.linalg.norm((a - b)[mask]) / max(np.linalg.norm(b[mask]), 1e-300)))
    # a short run through the reference-facing call: sample() on the REFERENCE density object
    tt = bfb.sample(den, dict(n_chain=64, n_iter=60, n_warmup=30, x_0=Xo[:24][np.arange(64) % 24], random_generator=3), verbose=False)
    h = dd._sync(False)
    out[name] = dict(e_logp=e_lp, e_grad=e_g, e_fit=e_fit, n_far=int(np.sum([den._surrogate_list[0]._use_bound])),
                     shape=list(tt.samples.shape), finite=bool(np.all(np.isfinite(tt.samples))),
                     original_ok=bool(np.allclose(tt.samples_original[3], den.to_original(tt.samples[3]))),
                     mean_depth=float(tt.arrays['tree_depth'].mean()))
print('RESULT ' + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason='baseline/_ref (the vendored reference) has not been built')
def test_real_reference_density_through_the_boundary():
    r = subprocess.run([sys.executable, '-c', SCRIPT, ROOT], capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')]
    assert r.returncode == 0 and line, r.stderr[-3000:]
    res = json.loads(line[-1][7:])
    assert set(res) == {'c2_n6_transform_decay', 'c2_n26_bound', 'c3_n5_decay'}
    for name, v in res.items():
        assert v['e_logp'] < 1e-10 and v['e_grad'] < 1e-10, (name, v)      # same relative metric as tests/test_gpu_eval.py
        assert v['e_fit'] < 1e-8, (name, v)                                # block-norm metric of tests/test_gpu_fit.py
        assert v['finite'] and v['original_ok'] and v['mean_depth'] >= 1., (name, v)
