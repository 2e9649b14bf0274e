"""GPU parity: batched PolyModel / Density evaluation through the C ABI vs golden vectors and the oracle."""
import pickle

import numpy as np
import pytest

import _golden_io as gio
from _specs import to_device_spec, synthetic_spec

pytestmark = pytest.mark.gpu
RTOL = 1e-10     # BASELINE.json north_star: values within 1e-10 relative error in FP64


def rel_err(a, b):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3 * np.max(np.abs(b)) + 1e-300))


@pytest.fixture(scope='module')
def handle():
    from bayesfast_b200 import _cabi
    h = _cabi.Handle(0)
    yield h
    h.close()


def test_device_rng_matches_spec(handle, oracle):
    u_d, z_d = handle.rng_fill(99, 12345678901, 5, 100000)
    u_h, z_h = oracle.rng_fill(99, 12345678901, 5, 100000)
    assert np.array_equal(u_d, u_h)                       # Philox + uniform conversion: bit exact
    central = np.abs(u_h - 0.5) <= 0.425                  # Phi^-1 central branch: fma-only arithmetic
    assert np.max(np.abs(z_d[central] - z_h[central]) / np.abs(z_h[central]).clip(1e-300)) < 4e-16
    assert np.max(np.abs(z_d - z_h)) < 1e-14              # tails go through log()/sqrt()


def test_poly_eval_golden(handle, oracle):
    for c in gio.load('poly_eval.npz')['cases']:
        handle.set_model(to_device_spec(c['spec']))
        F, J = handle.poly_eval_batch(c['X'])
        assert rel_err(F, c['wrapped_f']) < RTOL, c['name']
        assert rel_err(J, c['wrapped_j']) < RTOL, c['name']
        Fo, Jo = oracle.OracleDensity(c['spec']).poly_eval_batch(c['X'])
        assert rel_err(F, Fo) < RTOL and rel_err(J, Jo) < RTOL, c['name']


def test_poly_eval_c3n64_golden(handle, oracle):
    """BASELINE configs[3] at its named size (64-D cubic-3, P = 47905) against the real reference, inside and far outside the bound"""
    from _specs import c3n64_spec
    g = gio.load('poly_eval_c3n64.npz')
    handle.set_model(to_device_spec(c3n64_spec(g)))
    F, J = handle.poly_eval_batch(g['X'])
    assert rel_err(F, g['raw_f']) < RTOL and rel_err(J, g['raw_j']) < RTOL
    lp, gr = handle.logp_and_grad_batch(g['X'])
    assert rel_err(lp, g['raw_f'][:, 0]) < RTOL and rel_err(gr, g['raw_j'][:, 0]) < RTOL


def test_poly_kat(handle):
    g = gio.load('poly_kat.npz')
    c = g['logp']
    handle.set_model(to_device_spec(c['spec']))
    F, J = handle.poly_eval_batch(g['x'])
    assert rel_err(F, c['values']) < RTOL
    assert rel_err(J[0, 0], c['jac0']) < RTOL
    F, J = handle.poly_eval_batch(c['far'][None])
    assert rel_err(F[0], c['far_f']) < RTOL and rel_err(J[0], c['far_j']) < RTOL
    # SURVEY.md 8c known answers
    assert abs(F[0, 0] - (-46.7409242042374)) < 1e-9
    assert np.allclose(J[0, 0], [-11.740561821414351, 9.420776593640085, -18.24732501727469, -47.35933812148955],
                       rtol=1e-10)


def test_density_golden(handle, oracle):
    for c in gio.load('density.npz')['cases']:
        handle.set_model(to_device_spec(c['spec']))
        lp, g = handle.logp_and_grad_batch(c['X'])
        assert rel_err(lp, c['logp']) < RTOL, c['name']
        assert rel_err(g, c['grad']) < RTOL, c['name']


@pytest.mark.parametrize('n,order', [(26, 'cubic-2'), (16, 'cubic-2'), (2, 'quadratic'), (40, 'cubic-3'), (64, 'cubic-2')])
def test_density_large_batch_vs_oracle(handle, oracle, n, order):
    spec, cov = synthetic_spec(n, order, seed=n, decay=True, transform=(n == 26))
    handle.set_model(to_device_spec(spec))
    rng = np.random.default_rng(1)
    L = np.linalg.cholesky(cov)
    C = 4099 if n < 40 else 300                           # ragged: not a multiple of the block size
    X = (L @ rng.normal(size=(n, C))).T * rng.choice([0.5, 1., 3.], size=(C, 1))
    if spec['transform_ranges'] is not None:
        X = np.clip(X, spec['transform_ranges'][:, 0] * 0.9, spec['transform_ranges'][:, 1] * 0.9)
        X = np.array([oracle.from_original(x, spec['transform_ranges'], spec['hard_bounds']) for x in X])
    lp, g = handle.logp_and_grad_batch(X)
    lpo, go = oracle.OracleDensity(spec).logp_and_grad_batch(X)
    assert rel_err(lp, lpo) < RTOL and rel_err(g, go) < RTOL


@pytest.mark.parametrize('n,order,C', [(64, 'cubic-3', 1003), (40, 'cubic-3', 77), (48, 'cubic-2', 2500), (64, 'quadratic', 9), (33, 'cubic-2', 8)])
def test_team_evaluator_above_32_dimensions(handle, oracle, monkeypatch, n, order, C):
    """32 < n <= 64: logp + gradient on the tensor cores by teams of four warps (bfb_eval_team.cu; the cubic-3 pair-product operand
    streamed from L2) against the oracle and the generic kernel; points inside and far outside the radial bound, ragged counts"""
    spec, cov = synthetic_spec(n, order, seed=50 + n, cubic_scale=0.05)
    spec['alpha'] = spec['alpha'] / 1.6
    handle.set_model(to_device_spec(spec))
    rng = np.random.default_rng(2)
    X = (np.linalg.cholesky(cov) @ rng.normal(size=(n, C))).T * rng.choice([0.5, 1., 3.], size=(C, 1))
    lp, g = handle.logp_and_grad_batch(X)
    assert handle.eval_last_path() == 'team'
    lpo, go = oracle.OracleDensity(spec).logp_and_grad_batch(X)
    assert rel_err(lp, lpo) < RTOL and rel_err(g, go) < RTOL
    F, J = handle.poly_eval_batch(X)                     # PolyModel._fun_and_jac through the same evaluator
    Fo, Jo = oracle.OracleDensity(spec).poly_eval_batch(X)
    assert rel_err(F, Fo) < RTOL and rel_err(J, Jo) < RTOL
    monkeypatch.setenv('BFB200_EVAL', 'generic')
    lpg, gg = handle.logp_and_grad_batch(X)
    assert handle.eval_last_path() == 'generic'
    assert rel_err(lp, lpg) < RTOL and rel_err(g, gg) < RTOL


def test_empty_batch_and_errors(handle):
    spec, _ = synthetic_spec(5)
    handle.set_model(to_device_spec(spec))
    F, J = handle.poly_eval_batch(np.zeros((0, 5)))
    assert F.shape == (0, 1) and J.shape == (0, 1, 5)
    from bayesfast_b200 import _cabi
    bad = to_device_spec(spec)
    bad['configs'][0]['input_mask'] = np.array([0, 1, 2, 3, 7])
    with pytest.raises(_cabi.BfbError):
        handle.set_model(bad)


def test_polymodel_api(oracle):
    import bayesfast_b200 as bfb
    rng = np.random.default_rng(3)
    s = bfb.PolyModel('cubic-3', input_size=5, output_size=2, bound_options={'use_bound': False})
    for conf in s.configs:
        for i in range(conf.output_size):
            conf._set(rng.normal(size=conf._a_shape), i)
    x = rng.normal(size=5)
    f, j = s._fun_and_jac(x)
    spec = s.to_spec()
    Fo, Jo = oracle.OracleDensity(spec).poly_eval_batch(x[None])
    assert rel_err(f, Fo[0]) < RTOL and rel_err(j, Jo[0]) < RTOL
    assert isinstance(s(x), list) and s(x)[0].shape == (2,) and s.jac(x)[0].shape == (2, 5)
    s2 = pickle.loads(pickle.dumps(s))
    assert np.array_equal(s2._fun(x), s._fun(x))
    with pytest.raises(ValueError):
        s._fun(np.zeros(4))
    with pytest.raises(ValueError):
        bfb.PolyModel([bfb.PolyConfig('linear'), bfb.PolyConfig('linear')], input_size=2, output_size=1)


@pytest.mark.parametrize('n,order', [(26, 'cubic-2'), (28, 'cubic-2'), (16, 'cubic-2'), (5, 'cubic-2'), (31, 'cubic-2'),
                                     (26, 'quadratic'), (13, 'quadratic'), (32, 'quadratic')])
def test_tensor_core_evaluator_vs_oracle(handle, oracle, n, order, monkeypatch):
    """bfb_eval_dmma.cu (8 points per warp on FP64 DMMA) against the oracle, incl. points outside the radial bound
    (PolyModel._fj_bound) and a ragged batch; and against the generic warp-per-point kernel."""
    spec, cov = synthetic_spec(n, order, seed=100 + n)
    handle.set_model(to_device_spec(spec))
    rng = np.random.default_rng(2)
    L = np.linalg.cholesky(cov)
    C = 2051
    X = (L @ rng.normal(size=(n, C))).T * rng.choice([0.3, 1., 2.5, 6.], size=(C, 1))
    d = X - spec['mu']
    beta = np.einsum('ij,jk,ik->i', d, spec['hess'], d) ** 0.5
    assert (beta > spec['alpha']).sum() > 50 and (beta < spec['alpha']).sum() > 500
    lp, g = handle.logp_and_grad_batch(X)
    lpo, go = oracle.OracleDensity(spec).logp_and_grad_batch(X)
    assert rel_err(lp, lpo) < RTOL and rel_err(g, go) < RTOL
    monkeypatch.setenv('BFB200_EVAL', 'generic')
    lp2, g2 = handle.logp_and_grad_batch(X)
    assert rel_err(lp, lp2) < 1e-12 and rel_err(g, g2) < 1e-12
    monkeypatch.delenv('BFB200_EVAL')
    # batches smaller than a warp's 8 points
    lp3, g3 = handle.logp_and_grad_batch(X[:3])
    assert np.array_equal(lp3, lp[:3]) and np.array_equal(g3, g[:3])


def test_tensor_core_evaluator_extended_density(handle, oracle, monkeypatch):
    """decay + variable transform + module rescale through eval_dmma_kernel (model variant bit 1) vs the oracle and the
    generic kernel"""
    n = 26
    spec, cov = synthetic_spec(n, 'cubic-2', seed=77, decay=True, transform=True)
    rng = np.random.default_rng(4)
    s0 = -0.2 + 0.1 * rng.normal(size=n)
    spec['input_scales'] = np.stack((s0, s0 + 1.3 + 0.3 * rng.random(n)), axis=1)
    handle.set_model(to_device_spec(spec))
    C = 1500
    X = (np.linalg.cholesky(cov) @ rng.normal(size=(n, C))).T * rng.choice([0.5, 1., 2.5], size=(C, 1))
    X = np.clip(X, spec['transform_ranges'][:, 0] * 0.9, spec['transform_ranges'][:, 1] * 0.9)
    X = np.array([oracle.from_original(x, spec['transform_ranges'], spec['hard_bounds']) for x in X])
    lp, g = handle.logp_and_grad_batch(X)
    lpo, go = oracle.OracleDensity(spec).logp_and_grad_batch(X)
    assert rel_err(lp, lpo) < RTOL and rel_err(g, go) < RTOL
    monkeypatch.setenv('BFB200_EVAL', 'generic')
    lp2, g2 = handle.logp_and_grad_batch(X)
    assert rel_err(lp, lp2) < 1e-12 and rel_err(g, g2) < 1e-12


@pytest.mark.parametrize('n,kw', [(26, {}), (12, {}), (7, dict(decay=True, transform=True)), (28, {})])
def test_tensor_core_evaluator_cubic3(handle, oracle, monkeypatch, n, kw):
    """cubic-3 configs on the tensor-core evaluator (model variant bit 2: pair products x [pairs x n] coefficient GEMM)
    vs the oracle (_cubic_3_f / _cubic_3_j, _poly.pyx:86-137) and vs the generic kernel, inside and outside the bound"""
    spec, cov = synthetic_spec(n, 'cubic-3', seed=200 + n, cubic_scale=0.05, **kw)
    handle.set_model(to_device_spec(spec))
    rng = np.random.default_rng(3)
    C = 1037
    X = (np.linalg.cholesky(cov) @ rng.normal(size=(n, C))).T * rng.choice([0.4, 1., 2.5, 5.], size=(C, 1))
    if spec['transform_ranges'] is not None:
        X = np.clip(X, spec['transform_ranges'][:, 0] * 0.9, spec['transform_ranges'][:, 1] * 0.9)
        X = np.array([oracle.from_original(x, spec['transform_ranges'], spec['hard_bounds']) for x in X])
    lp, g = handle.logp_and_grad_batch(X)
    lpo, go = oracle.OracleDensity(spec).logp_and_grad_batch(X)
    assert rel_err(lp, lpo) < RTOL and rel_err(g, go) < RTOL
    monkeypatch.setenv('BFB200_EVAL', 'generic')
    lp2, g2 = handle.logp_and_grad_batch(X)
    assert rel_err(lp, lp2) < 1e-11 and rel_err(g, g2) < 1e-11
