"""Host-side logic of the reference-interface mirrors (no device needed)."""
import pickle
from types import SimpleNamespace

import numpy as np
import pytest

import bayesfast_b200 as bfb
from bayesfast_b200 import transforms as tf
from bayesfast_b200.poly import pack_dense, unpack_dense
from bayesfast_b200.runtime import shard_bounds
from bayesfast_b200.sample_trace import NTrace, HTrace, TraceTuple, _get_step_size


def test_polyconfig_and_recipe():
    s = bfb.PolyModel('cubic-3', input_size=4, output_size=1)
    assert [c.order for c in s.configs] == ['linear', 'quadratic', 'cubic-2', 'cubic-3']
    assert s.n_param == 5 + 10 + 16 + 4                 # reference: poly.py:592-593
    assert bfb.PolyModel('cubic-2', input_size=26, output_size=1).n_param == 1054
    assert np.array_equal(s.recipe, [[0, 1, 2, 3]])
    with pytest.raises(ValueError):
        bfb.PolyConfig('quartic')
    with pytest.raises(ValueError):
        bfb.PolyModel([bfb.PolyConfig('quadratic', output_mask=[0]), bfb.PolyConfig('quadratic', output_mask=[0, 1])],
                      input_size=3, output_size=2)
    with pytest.raises(ValueError):
        bfb.PolyModel([bfb.PolyConfig('linear', output_mask=[0])], input_size=3, output_size=2)   # output 1 uncovered
    c = bfb.PolyConfig('quadratic', input_mask=[3, 1, 1], output_mask=[0])
    assert np.array_equal(c.input_mask, [1, 3]) and c._a_shape == (3,) and c._A_shape == (1, 2, 2)
    with pytest.raises(ValueError):
        c._set(np.zeros(4), 0)
    c._set(np.array([1., 2., 3.]), 0)
    assert np.array_equal(c._coef[0], [[1., 2.], [0., 3.]])
    with pytest.raises(ValueError):
        s.set_bound_options(alpha=None, alpha_p=None)


def test_pack_unpack_roundtrip(oracle):
    rng = np.random.default_rng(0)
    for order, n in (('linear', 5), ('quadratic', 6), ('cubic-2', 4), ('cubic-3', 7)):
        a = rng.normal(size=bfb.PolyConfig(order, np.arange(n), [0])._a_shape)
        dense = unpack_dense(order, a, n)
        assert np.array_equal(dense, oracle.unpack_coef(order, a, n))
        assert np.array_equal(pack_dense(order, dense, n), a)


def test_transforms_match_oracle_and_reference_test(oracle):
    # the configuration of the reference's tests/test_constraint.py:5-7
    ranges = np.array([[-10, 10], [-5, 8], [-4, 6], [-8, 6]], float)
    hb = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], np.uint8)
    x = np.ones(4) * 0.6
    f, j, jj = oracle.to_original(x, ranges, hb)
    assert np.allclose(tf.to_original(x, ranges, hb), f, rtol=1e-15)
    assert np.allclose(tf.to_original_grad(x, ranges, hb), j, rtol=1e-15)
    assert np.allclose(tf.to_original_grad2(x, ranges, hb), jj, rtol=1e-14)
    eps = 1e-6                                           # analytic vs numerical, like the reference test
    num = (tf.to_original(x + eps, ranges, hb) - tf.to_original(x - eps, ranges, hb)) / (2 * eps)
    assert np.allclose(j, num, rtol=1e-7)
    num2 = (tf.to_original_grad(x + eps, ranges, hb) - tf.to_original_grad(x - eps, ranges, hb)) / (2 * eps)
    assert np.allclose(jj, num2, rtol=1e-6, atol=1e-8)
    xo = tf.to_original(x, ranges, hb)
    assert np.allclose(tf.from_original(xo, ranges, hb), x, rtol=1e-12)
    assert np.allclose(oracle.from_original(xo, ranges, hb), x, rtol=1e-12)
    assert np.allclose(tf.from_original_grad(xo, ranges, hb) * j, 1., rtol=1e-12)
    with pytest.raises(ValueError):
        tf.from_original(np.array([0., 0., 0., 7.]), ranges, hb)      # variable #3 out of bound
    X = np.stack([x, 0.5 * x])
    assert tf.to_original(X, ranges, hb).shape == (2, 4)


def test_density_host_logic_and_pickle():
    s = bfb.PolyModel('quadratic', input_size=3, output_size=1, input_scales=[[0, 2], [0, 2], [-1, 1]])
    for c in s.configs:
        c._set(np.arange(c._a_shape[0], dtype=float), 0)
    d = bfb.Density(s, input_scales=[[-1, 1]] * 3, hard_bounds=[[1, 1], [0, 0], [1, 0]],
                    decay_options=dict(use_decay=True, alpha=3., alpha_p=None, gamma=0.2))
    d._mu, d._hess = np.zeros(3), np.eye(3)
    spec = d.to_spec()
    assert spec['use_decay'] and spec['d_alpha2'] == 9. and spec['d_gamma'] == 0.2
    assert spec['transform_ranges'].shape == (3, 2) and spec['hard_bounds'].tolist() == [[1, 1], [0, 0], [1, 0]]
    assert spec['input_scales'].shape == (3, 2) and not spec['use_bound']
    x = np.array([0.1, -0.3, 0.2])
    assert np.allclose(d.from_original(d.to_original(x)), x)
    assert np.isclose(d.to_original_density(1., x_trans=x), 1. - np.sum(np.log(np.abs(d.to_original_grad(x)))))
    d2 = pickle.loads(pickle.dumps(d))
    assert d2._handle is None and np.array_equal(d2.surrogate.configs[1]._coef, s.configs[1]._coef)
    with pytest.raises(ValueError):
        bfb.Density(object())
    with pytest.raises(ValueError):
        d.set_decay_options(gamma=-1.)


def test_from_reference_duck_typing():
    """a stand-in with the attribute layout of a fitted reference Density / PolyModel (core/density.py, modules/poly.py)"""
    n = 3
    lin = SimpleNamespace(order='linear', _input_mask=np.arange(n), _output_mask=np.arange(1), _coef=np.array([[1., 2., 3., 4.]]))
    quad = SimpleNamespace(order='quadratic', _input_mask=np.arange(n), _output_mask=np.arange(1),
                           _coef=np.triu(np.arange(9.).reshape(1, 3, 3)[0])[None])
    sur = SimpleNamespace(_configs=(lin, quad), _recipe=None, _scope=(0, 1), _use_bound=True, _alpha=2.5, _alpha_p=100.,
                          _center_max=True, _input_size=n, _output_size=1, _input_scales=None, _mu=np.zeros(n),
                          _hess=np.eye(n), _f_mu=np.array([0.5]))
    ref = SimpleNamespace(_surrogate_list=[sur], _module_list=[object()], use_surrogate=True, _input_scales=None,
                          _hard_bounds=False, _use_decay=False, _alpha=None, _alpha_p=150., _gamma=0.1, density_name='logp')
    d = bfb.Density.from_reference(ref)
    spec = d.to_spec()
    assert spec['use_bound'] and spec['alpha'] == 2.5 and np.array_equal(spec['configs'][1]['packed'][0], [0, 1, 2, 4, 5, 8])
    ref2 = SimpleNamespace(**{**ref.__dict__, '_module_list': [object(), object()]})
    with pytest.raises(ValueError, match='whole module list'):
        bfb.Density.from_reference(ref2)
    with pytest.raises(ValueError):
        bfb.sample(object())


def test_trace_objects():
    with pytest.raises(ValueError):
        NTrace(n_iter=10, n_warmup=10)
    # dense mass matrix: 'full', a covariance or a QuadMetricFull (sample_trace.py:375-390, 424-455)
    from bayesfast_b200.sample_trace import QuadMetricFull, QuadMetricFullAdapt, _get_metric
    assert NTrace(metric='full').metric == 'full' and NTrace(metric=np.eye(3)).metric.shape == (3, 3)
    with pytest.raises(ValueError):
        NTrace(metric=np.ones((2, 3)))
    with pytest.raises(ValueError, match='positive definite'):
        QuadMetricFull(np.diag([1., -1.]))
    with pytest.raises(ValueError):
        NTrace(max_treedepth=0)
    t = NTrace(n_chain=3, n_iter=6, n_warmup=2, x_0=np.zeros((3, 2)), random_generator=7)
    cfg = t._cfg_dict(7, 5)
    assert cfg['max_treedepth'] == 10 and cfg['chain0'] == 5 and cfg['n_warmup'] == 2 and cfg['target_accept'] == 0.8
    assert HTrace(n_int_step=8).n_call == 1500 * 9 + 1
    C, n_it, n = 3, 6, 2
    rng = np.random.default_rng(0)
    arrays = dict(samples=rng.normal(size=(C, n_it, n)), logp=rng.normal(size=(C, n_it)))
    arrays['samples_original'], arrays['logp_original'] = arrays['samples'], arrays['logp']
    for k in ('energy', 'mean_tree_accept', 'step_size', 'step_size_bar', 'energy_change', 'max_energy_change'):
        arrays[k] = rng.normal(size=(C, n_it))
    arrays['tree_depth'] = np.full((C, n_it), 2, np.int32)
    arrays['tree_size'] = np.full((C, n_it), 3, np.int32)
    arrays['diverging'] = np.zeros((C, n_it), np.int32)
    final = dict(final_step=np.tile([np.log(0.3), np.log(0.25), 0.1, 4.], (C, 1)), final_var=np.ones((C, n)),
                 step0=np.full(C, 0.5), x_0=np.zeros((C, n)))
    tt = TraceTuple(t, arrays, final, chain0=10)
    assert len(tt) == 3 and tt.sampler == 'NUTS' and tt.get().shape == (C * 4, n) and tt.get(return_type='logp').shape == (12,)
    t1 = tt[1]
    assert t1.chain_id == 11 and t1.samples.shape == (n_it, n) and t1.stats.n_warmup == 2 and t1.i_iter == 6
    assert t1.n_call == 3 * 5 + 6 + 1 and tt.n_call == 3 * (15 + 7)
    assert np.isclose(t1.step_size.current(False), 0.25) and np.isclose(_get_step_size(tt), 0.25 * 2**0.25)
    assert list(t1.stats.get().keys())[2] == 'tree_depth' and len(t1.stats.get()['logp']) == 4
    with pytest.raises(ValueError):
        tt.get(since_iter=5)
    # a dense-metric run hands back covariances: the per-chain traces carry QuadMetricFullAdapt, _get_metric averages them
    cov = np.array([[2., 0.3], [0.3, 1.]])
    ttd = TraceTuple(t, arrays, dict(final, final_var=np.tile(cov, (C, 1, 1)), chol_error=np.array([0, 1, 0], np.int32)))
    assert isinstance(ttd[0].metric, QuadMetricFullAdapt) and np.array_equal(ttd[0].metric._cov, cov)
    ttd[0].metric.raise_ok()
    with pytest.raises(ValueError, match='Cholesky'):
        ttd[1].metric.raise_ok()
    assert np.allclose(_get_metric(ttd, 'full', from_samples=False), cov)
    assert np.allclose(_get_metric(ttd[2], 'diag', from_samples=False), [2., 1.])


def test_shard_bounds():
    for total, world in ((4096, 8), (10, 3), (2, 2), (7, 8)):
        cover = []
        for r in range(world):
            lo, hi = shard_bounds(total, r, world)
            cover += list(range(lo, hi))
        assert cover == list(range(total))


def test_likelihood_whitening_matches_oracle(oracle):
    """whiten_spec folds the Gaussian likelihood into the polynomial coefficients (f' = Lt (f - d)); the oracle evaluates the
    un-whitened pipeline the way the reference does: both must agree (golden pipelines of the real reference)."""
    import _golden_io as gio
    from _specs import to_device_spec
    from bayesfast_b200.density import whiten_spec, GaussianLikelihood
    from bayesfast_b200.poly import unpack_dense
    for c in gio.load('pipeline.npz')['cases']:
        spec = to_device_spec(c['spec'])
        ep = spec['epilogue']
        lik = GaussianLikelihood(ep['d'], ep['cinv'], ep['c0'])
        w = whiten_spec(spec, lik)
        assert w['epilogue_sumsq'] == ep['c0'] and 'epilogue' not in w
        assert all(len(cf['input_mask']) == spec['n'] and len(cf['output_mask']) == spec['m'] for cf in w['configs'])
        for cf in w['configs']:
            cf['coef'] = np.array([unpack_dense(cf['order'], a, int(spec['n'])) for a in cf['packed']])
        F, J = oracle.OracleDensity(dict(w, use_decay=False, transform_ranges=None)).poly_eval_batch(c['X'])
        lp0, gr0 = oracle.OracleDensity(dict(c['spec'], use_decay=False)).logp_and_grad_batch(c['X'])
        assert np.allclose(lik.const - 0.5 * np.sum(F * F, axis=1), lp0, rtol=1e-12, atol=1e-12)
        assert np.allclose(-np.einsum('co,con->cn', F, J), gr0, rtol=1e-12, atol=1e-12 * np.abs(gr0).max())
        assert np.allclose(lik.logp(np.zeros(spec['m'])), ep['c0'] - 0.5 * ep['d'] @ ep['cinv'] @ ep['d'])
    with pytest.raises(ValueError, match='semi-definite'):
        GaussianLikelihood([0., 0.], [[1., 0.], [0., -1.]])
    sur = bfb.PolyModel('quadratic', input_size=2, output_size=1)
    with pytest.raises(ValueError, match='outputs'):
        bfb.Density(sur, likelihood=GaussianLikelihood([0., 0.], np.eye(2)))


def test_from_reference_with_likelihood():
    """a reference Density [module_0 (x -> m), module_1 (m -> logp)] whose surrogate replaces module_0 only (scope (0, 1)) is
    accepted when the caller restates module_1 as a GaussianLikelihood, and refused otherwise"""
    from types import SimpleNamespace
    n = 2
    conf = SimpleNamespace(order='quadratic', _input_mask=np.arange(n), _output_mask=np.arange(1),
                           _coef=np.array([[[0.5, 0.1], [0., 0.25]]]))
    lin = SimpleNamespace(order='linear', _input_mask=np.arange(n), _output_mask=np.arange(1), _coef=np.array([[1., 0.2, -0.3]]))
    rs = SimpleNamespace(_configs=[lin, conf], _recipe=None, _scope=(0, 1), _use_bound=False, _alpha=None, _alpha_p=100.,
                         _center_max=True, _input_size=n, _output_size=1, _input_scales=None)
    ref = SimpleNamespace(_surrogate_list=[rs], _module_list=[object(), object()], use_surrogate=True, _input_scales=None,
                          _hard_bounds=False, _use_decay=False, _alpha=None, _alpha_p=150., _gamma=0.1, density_name='logp')
    with pytest.raises(ValueError, match='whole module list'):
        bfb.Density.from_reference(ref)
    lik = bfb.GaussianLikelihood([5.], [[4.]])
    den = bfb.Density.from_reference(ref, likelihood=lik)
    spec = den.to_spec()
    assert spec['epilogue']['cinv'][0, 0] == 4. and spec['m'] == 1 and den.likelihood is lik
    from bayesfast_b200.density import whiten_spec
    w = whiten_spec(spec, lik)
    # f' = 2 (f - 5): constant 2 (1 - 5), linear and quadratic coefficients doubled
    assert np.allclose(w['configs'][0]['packed'], [[-8., 0.4, -0.6]]) and np.allclose(w['configs'][1]['packed'], [[1., 0.2, 0.5]])
    rs._scope = (0, 2)
    with pytest.raises(ValueError, match='all modules but the last'):
        bfb.Density.from_reference(ref, likelihood=lik)
    import pickle
    den2 = pickle.loads(pickle.dumps(den))
    assert den2.likelihood.inv_cov[0, 0] == 4. and den2._handle is None


def test_likelihood_whitening_all_orders_and_masks(oracle):
    """whiten_spec with masked cubic-2 / cubic-3 configs and overlapping output masks (two configs feeding the same output)"""
    from _specs import to_device_spec
    from bayesfast_b200.density import whiten_spec, GaussianLikelihood
    from bayesfast_b200.poly import unpack_dense
    rng = np.random.default_rng(12)
    n, m = 6, 4
    c3 = np.zeros((2, 4, 4, 4))
    for j in range(4):
        for k in range(j + 1, 4):
            c3[:, j, k, k + 1:] = rng.normal(size=(2, 4 - k - 1)) * 0.1
    cfgs = [dict(order='linear', input_mask=np.array([0, 2, 5]), output_mask=np.arange(m), coef=rng.normal(size=(m, 4))),
            dict(order='quadratic', input_mask=np.array([1, 2, 4]), output_mask=np.array([0, 3]), coef=np.triu(rng.normal(size=(2, 3, 3)))),
            dict(order='quadratic', input_mask=np.arange(n), output_mask=np.array([3]), coef=np.triu(rng.normal(size=(1, n, n))) * 0.3),
            dict(order='cubic-2', input_mask=np.array([0, 3]), output_mask=np.array([1, 2]), coef=rng.normal(size=(2, 2, 2)) * 0.2),
            dict(order='cubic-3', input_mask=np.array([0, 1, 3, 5]), output_mask=np.array([0, 2]), coef=c3)]
    B = rng.normal(size=(m, m))
    ep = dict(d=rng.normal(size=m), cinv=B @ B.T + 0.1 * np.eye(m), c0=2.5)
    spec = dict(n=n, m=m, configs=cfgs, use_bound=True, mu=rng.normal(size=n) * 0.1, hess=np.eye(n) * 0.5, alpha=1.2,
                f_mu=rng.normal(size=m), input_scales=None, use_decay=False, transform_ranges=None, epilogue=ep)
    lik = GaussianLikelihood(ep['d'], ep['cinv'], ep['c0'])
    w = whiten_spec(to_device_spec(spec), lik)
    assert [c['order'] for c in w['configs']] == ['linear', 'quadratic', 'cubic-2', 'cubic-3']
    for cf in w['configs']:
        cf['coef'] = np.array([unpack_dense(cf['order'], a, n) for a in cf['packed']])
    X = np.concatenate((rng.normal(size=(12, n)) * 0.5, rng.normal(size=(8, n)) * 3.))      # inside and outside the bound
    F, J = oracle.OracleDensity(w).poly_eval_batch(X)
    lp0, gr0 = oracle.OracleDensity(spec).logp_and_grad_batch(X)
    assert np.allclose(lik.const - 0.5 * np.sum(F * F, axis=1), lp0, rtol=1e-11, atol=1e-11)
    assert np.allclose(-np.einsum('co,con->cn', F, J), gr0, rtol=1e-11, atol=1e-11 * np.abs(gr0).max())


def test_default_x0_is_the_reference_sobol_sequence():
    """sample() without x_0 starts the chains at the reference's Sobol multivariate-normal points (core/sample.py:107-112,
    utils/sobol.py:48-60): bit-identical to utils.sobol.multivariate_normal of the real reference (tests/golden/sobol_x0.npz)"""
    import os
    from bayesfast_b200.sample import sobol_multivariate_normal
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'sobol_x0.npz'))
    for k in g.files:
        d, n = int(k.split('_')[0][1:]), int(k.split('_')[1][1:])
        assert np.array_equal(sobol_multivariate_normal(d, n), g[k]), k


def test_tempered_trace_objects():
    """TNTrace / THTrace (samplers/sample_trace.py:540-622), TNStats / THStats (hmc_utils/stats.py:9-24, 103-118): host mirrors"""
    from bayesfast_b200.sample_trace import TNTrace, THTrace, TNStats, THStats, tnstats_items, thstats_items
    base = SimpleNamespace(logp_and_grad=lambda x: (0., x))
    with pytest.raises(ValueError, match='density_base'):
        TNTrace(object())
    with pytest.raises(ValueError, match='logxi'):
        TNTrace(base, logxi='a')
    t = TNTrace(base, 0.3, n_chain=3, n_iter=6, n_warmup=2, x_0=np.zeros((3, 2)), random_generator=7, u_0=[0.1, 0.2, 0.3],
                max_treedepth=7)
    assert t.logxi == 0.3 and t.density_base is base and np.array_equal(t.u_0, [0.1, 0.2, 0.3]) and t.max_treedepth == 7
    assert isinstance(t, NTrace) and isinstance(t.stats, TNStats) and t.stats.stats_items == tnstats_items
    assert tnstats_items[:2] == ('u', 'weight') and thstats_items[:3] == ('u', 'weight', 'logp')
    th = THTrace(base, n_int_step=5, n_chain=2)
    assert isinstance(th, HTrace) and isinstance(th.stats, THStats) and th.n_int_step == 5 and th.u_0 is None and th.logxi == 0.
    assert t._cfg_dict(1, 0)['max_treedepth'] == 7 and th._cfg_dict(1, 0)['n_int_step'] == 5
    C, n_it, n = 3, 6, 2
    rng = np.random.default_rng(0)
    arrays = dict(samples=rng.normal(size=(C, n_it, n)), logp=rng.normal(size=(C, n_it)), u=rng.normal(size=(C, n_it)),
                  weight=rng.uniform(size=(C, n_it)))
    arrays['samples_original'], arrays['logp_original'] = arrays['samples'], arrays['logp']
    for k in ('energy', 'mean_tree_accept', 'step_size', 'step_size_bar', 'energy_change', 'max_energy_change'):
        arrays[k] = rng.normal(size=(C, n_it))
    arrays['tree_depth'] = np.full((C, n_it), 2, np.int32)
    arrays['tree_size'] = np.full((C, n_it), 3, np.int32)
    arrays['diverging'] = np.zeros((C, n_it), np.int32)
    final = dict(final_step=np.tile([np.log(0.3), np.log(0.25), 0.1, 4.], (C, 1)), final_var=np.ones((C, n)),
                 step0=np.full(C, 0.5), x_0=np.zeros((C, n)))
    tt = TraceTuple(t, arrays, final)
    assert tt.sampler == 'TNUTS' and tt.n_call == 3 * (15 + 7)
    t1 = tt[1]
    assert isinstance(t1, TNTrace) and isinstance(t1.stats, TNStats)
    assert np.array_equal(t1.stats._u, arrays['u'][1]) and np.array_equal(t1.stats.weight, arrays['weight'][1])
    assert np.array_equal(t1.get(return_type='weights'), arrays['weight'][1][2:])            # sample_trace.py:574-585
    s, lp, w = t1.get(return_type='all')
    assert s.shape == (4, n) and lp.shape == (4,) and w.shape == (4,)
    tth = TraceTuple(th, dict(arrays), dict(final))
    assert tth.sampler == 'THMC' and isinstance(tth[0].stats, THStats) and tth[0].stats._accepted.dtype == bool
    assert np.array_equal(tth[0].stats._n_int_step, arrays['tree_size'][0])
    # sample(): the sampler names are accepted, the trace needs its base density
    with pytest.raises(TypeError):
        bfb.sample(bfb.Density(bfb.PolyModel('quadratic', input_size=2, output_size=1)), {}, sampler='TNUTS')
