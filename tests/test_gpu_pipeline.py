"""GPU parity of the two-module pipeline (PolyModel surrogate -> Gaussian-likelihood module; core/density.py:487-566):
BASELINE configs[0] (2-D donut of examples/2d-donut.ipynb) and a multi-output surrogate with masked configs, against recorded
runs of the real reference (tests/golden/pipeline.npz) and against the oracle."""
import numpy as np
import pytest

import _golden_io as gio
from _specs import to_device_spec
from test_gpu_sampler import cfg_from, device_draws, check_floats, INT_STATS, FLT_STATS

pytestmark = pytest.mark.gpu
CASES = gio.load('pipeline.npz')['cases']


@pytest.fixture(scope='module')
def handle():
    from bayesfast_b200 import _cabi
    h = _cabi.Handle(0)
    yield h
    h.close()


def _whitened(case):
    from bayesfast_b200.density import whiten_spec, GaussianLikelihood
    spec = to_device_spec(case['spec'])
    ep = spec['epilogue']
    return whiten_spec(spec, GaussianLikelihood(ep['d'], ep['cinv'], ep['c0']))


@pytest.mark.parametrize('case', CASES, ids=lambda c: c['name'])
def test_pipeline_logp_and_grad_golden(handle, case):
    handle.set_model(_whitened(case))
    lp, g = handle.logp_and_grad_batch(case['X'])
    assert np.allclose(lp, case['logp'], rtol=1e-10, atol=1e-10)
    assert np.allclose(g, case['grad'], rtol=1e-10, atol=1e-10 * np.abs(case['grad']).max())


@pytest.mark.parametrize('case', CASES, ids=lambda c: c['name'])
def test_pipeline_golden_chains(handle, case):
    """NUTS through the pipeline: same seeds and draw stream as the recorded runs of the real reference"""
    r, kw = case['result'], case['trace_kw']
    n_iter, n_warmup = int(kw['n_iter']), int(kw['n_warmup'])
    handle.set_model(_whitened(case))
    handle.sampler_init(cfg_from(kw, n_warmup, int(case['seed'])), case['x0'], float(r['step0']), r['var0'], case['x0'])
    out = handle.sampler_run('NUTS', n_iter)
    assert handle.sampler_last_path() == 'generic'
    st = handle.sampler_state()
    assert np.all(st['status'] == 0)
    assert np.array_equal(st['n_draws'], r['n_draws'])
    for k in INT_STATS:
        assert np.array_equal(out[k], r[k].astype(np.int32)), k
    # first trees of the donut are 6-7 doublings deep: 1e-9 over the first 4 iterations (see tests/test_oracle_golden.py)
    for k in FLT_STATS:
        assert np.allclose(out[k][:, :4], r[k][:, :4], rtol=1e-9, atol=1e-9), k
        assert np.allclose(out[k], r[k], rtol=1e-3, atol=1e-3), k
    assert np.allclose(out['samples'], r['samples'], rtol=1e-3, atol=1e-3)


def test_donut_public_api(oracle):
    """BASELINE configs[0] through the public interface: quadratic PolyModel surrogate of m = |x| fitted on the device,
    GaussianLikelihood([5], [[4]]) = f_1 of the notebook, decay on, bound off, 4 NUTS chains; the oracle restates the pipeline"""
    import bayesfast_b200 as bfb
    rng = np.random.default_rng(3)
    th, rad = rng.uniform(0., 2. * np.pi, size=30), 5. + 0.5 * rng.normal(size=30)
    xf = np.stack((rad * np.cos(th), rad * np.sin(th)), axis=1)
    sur = bfb.PolyModel('quadratic', input_size=2, output_size=1, bound_options=dict(use_bound=False))
    den = bfb.Density(sur, decay_options=dict(use_decay=True), likelihood=bfb.GaussianLikelihood([5.], [[4.]]))
    den.fit(xf, np.linalg.norm(xf, axis=1))
    spec = den.to_spec()
    assert spec['epilogue']['cinv'][0, 0] == 4. and spec['use_decay'] and not spec['use_bound']
    od = oracle.OracleDensity(spec)
    X = np.concatenate((xf[:8] * 1.03, rng.normal(size=(6, 2)) * 2., xf[8:12] * 3.))
    lp, g = den.logp_and_grad(X, original_space=False)
    lpo, go = od.logp_and_grad_batch(X)
    assert np.allclose(lp, lpo, rtol=1e-10, atol=1e-10) and np.allclose(g, go, rtol=1e-10, atol=1e-10 * np.abs(go).max())
    # a quadratic in x approximates |x| on the ring only roughly: the pipeline's logp follows the true -(|x| - 5)^2 / 0.5
    assert np.max(np.abs(lp[:8] + (np.linalg.norm(X[:8], axis=1) - 5.)**2 / 0.5)) < 1.
    x0 = xf[12:16].copy()
    tt = bfb.sample(den, dict(n_chain=4, n_iter=60, n_warmup=30, x_0=x0, random_generator=11), verbose=False)
    assert tt.samples.shape == (4, 60, 2)
    U, Z = device_draws(den._sync(False), 11, tt._final['n_draws'])
    ref = od.run('NUTS', dict(n_iter=60, n_warmup=30), x0, 1. / 2**0.25, np.ones(2), draws_u=U, draws_z=Z)
    for k in INT_STATS:
        assert np.array_equal(tt.arrays[k], ref[k]), k
    assert np.allclose(tt.samples, ref['samples'], rtol=1e-3, atol=1e-3)
    rr = np.linalg.norm(tt.samples[:, 30:].reshape(-1, 2), axis=1)
    assert abs(rr.mean() - 5.) < 1.


def test_multi_output_likelihood_vs_oracle(handle, oracle):
    """m = 40 outputs (block-quadratic masked configs), n = 12: evaluation and teacher-forced NUTS against the oracle"""
    from bayesfast_b200.density import whiten_spec, GaussianLikelihood
    rng = np.random.default_rng(8)
    n, m, C, n_iter = 12, 40, 48, 30
    masks = [np.sort(rng.choice(n, size=5, replace=False)) for _ in range(4)]
    cfgs = [dict(order='linear', input_mask=np.arange(n), output_mask=np.arange(m),
                 coef=np.concatenate((rng.normal(size=(m, 1)) * 0.2, rng.normal(size=(m, n)) * 0.5), axis=1))]
    for b, im in enumerate(masks):
        q = np.triu(rng.normal(size=(10, 5, 5))) * 0.08
        cfgs.append(dict(order='quadratic', input_mask=im, output_mask=np.arange(10 * b, 10 * b + 10), coef=q))
    B = rng.normal(size=(m, m))
    ep = dict(d=rng.normal(size=m) * 0.2, cinv=B @ B.T / m + 0.3 * np.eye(m), c0=0.75)
    spec = dict(n=n, m=m, configs=cfgs, use_bound=False, input_scales=None, use_decay=False, transform_ranges=None, epilogue=ep)
    w = whiten_spec(to_device_spec(spec), GaussianLikelihood(ep['d'], ep['cinv'], ep['c0']))
    handle.set_model(w)
    X = rng.normal(size=(64, n)) * 0.7
    lp, g = handle.logp_and_grad_batch(X)
    od = oracle.OracleDensity(spec)
    lpo, go = od.logp_and_grad_batch(X)
    assert np.allclose(lp, lpo, rtol=1e-10, atol=1e-10) and np.allclose(g, go, rtol=1e-10, atol=1e-10 * np.abs(go).max())
    x0 = rng.normal(size=(C, n)) * 0.3
    seed = 515
    handle.sampler_init(cfg_from({}, n_iter // 2, seed), x0, 1. / n**0.25, np.ones(n), x0)
    out = handle.sampler_run('NUTS', n_iter)
    st = handle.sampler_state()
    assert np.all(st['status'] == 0)
    U, Z = device_draws(handle, seed, st['n_draws'])
    ref = od.run('NUTS', dict(n_iter=n_iter, n_warmup=n_iter // 2), x0, 1. / n**0.25, np.ones(n), draws_u=U, draws_z=Z)
    assert np.array_equal(st['n_draws'], ref['n_draws'])
    for k in INT_STATS:
        assert np.array_equal(out[k], ref[k]), k
    check_floats(out['samples'], ref['samples'], 'samples')


def _lik_spec(rng, n, m, with_bound=False):
    """block-quadratic masked surrogate (DES-Y1 shape) + dense inverse covariance"""
    nb = max(1, m // 10)
    edges = np.linspace(0, m, nb + 1).astype(int)
    cfgs = [dict(order='linear', input_mask=np.arange(n), output_mask=np.arange(m),
                 coef=np.concatenate((rng.normal(size=(m, 1)) * 0.2, rng.normal(size=(m, n)) * 0.4), axis=1))]
    for b in range(nb):
        k, ni = edges[b + 1] - edges[b], min(n, 6)
        cfgs.append(dict(order='quadratic', input_mask=np.sort(rng.choice(n, size=ni, replace=False)),
                         output_mask=np.arange(edges[b], edges[b + 1]), coef=np.triu(rng.normal(size=(k, ni, ni))) * 0.1))
    B = rng.normal(size=(m, m))
    ep = dict(d=rng.normal(size=m) * 0.2, cinv=B @ B.T / m + 0.3 * np.eye(m), c0=-0.5)
    spec = dict(n=n, m=m, configs=cfgs, use_bound=False, input_scales=None, use_decay=False, transform_ranges=None, epilogue=ep)
    if with_bound:
        spec.update(use_bound=True, mu=np.zeros(n), hess=np.eye(n), alpha=1.5, f_mu=rng.normal(size=m))
    return spec


@pytest.mark.parametrize('n,m,C', [(26, 61, 1000), (12, 40, 7), (30, 9, 129), (16, 33, 1), (5, 3, 64)])
def test_likelihood_tensor_core_evaluator(handle, oracle, monkeypatch, n, m, C):
    """lik_eval_dmma_kernel (bfb_lik_dmma.cu: the outputs' S_o x as one DMMA GEMM per 8 points, operand chunks staged in
    shared memory) against the oracle and against the generic kernel, ragged point counts included"""
    from bayesfast_b200.density import whiten_spec, GaussianLikelihood
    rng = np.random.default_rng(100 + n + m)
    spec = _lik_spec(rng, n, m)
    ep = spec['epilogue']
    w = whiten_spec(to_device_spec(spec), GaussianLikelihood(ep['d'], ep['cinv'], ep['c0']))
    X = rng.normal(size=(C, n)) * 0.6
    lpo, go = oracle.OracleDensity(spec).logp_and_grad_batch(X)
    handle.set_model(w)
    lp, g = handle.logp_and_grad_batch(X)
    assert handle.eval_last_path() in ('lik_dmma', 'lik_feat')      # the feature form when the outputs' masks have a small union
    monkeypatch.setenv('BFB200_EVAL', 'generic')
    lpg, gg = handle.logp_and_grad_batch(X)
    assert handle.eval_last_path() == 'generic'
    monkeypatch.delenv('BFB200_EVAL')
    for a, b in ((lp, lpo), (lpg, lpo)):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-10)
    for a, b in ((g, go), (gg, go)):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-10 * np.abs(go).max())
    # small staging chunks (several chunks per pass, a ragged last one) give the same bits
    monkeypatch.setenv('BFB200_LIK_CHUNK', '4')
    lp2, g2 = handle.logp_and_grad_batch(X)
    assert np.array_equal(lp2, lp) and np.array_equal(g2, g)


def test_likelihood_with_bound_on_both_evaluators(handle, oracle, monkeypatch):
    """a radial bound (poly.py:466-503) applies to every output: the tensor-core evaluator (bound term applied once to the
    accumulated gradient, bfb_dmma.cuh lik_post) and the generic one both match the oracle"""
    from bayesfast_b200.density import whiten_spec, GaussianLikelihood
    rng = np.random.default_rng(4)
    spec = _lik_spec(rng, 8, 12, with_bound=True)
    ep = spec['epilogue']
    handle.set_model(whiten_spec(to_device_spec(spec), GaussianLikelihood(ep['d'], ep['cinv'], ep['c0'])))
    X = rng.normal(size=(40, 8)) * 0.9              # radius 1.5: a good part of the points is outside
    lpo, go = oracle.OracleDensity(spec).logp_and_grad_batch(X)
    assert np.sum(np.linalg.norm(X, axis=1) > 1.5) > 5
    for path in ('tensor', 'generic'):
        if path == 'generic':
            monkeypatch.setenv('BFB200_EVAL', 'generic')
        lp, g = handle.logp_and_grad_batch(X)
        assert handle.eval_last_path() == ('generic' if path == 'generic' else 'lik_feat')       # n = 8: P_f = 45 features
        assert np.allclose(lp, lpo, rtol=1e-10, atol=1e-10) and np.allclose(g, go, rtol=1e-10, atol=1e-10 * np.abs(go).max())


@pytest.mark.parametrize('n,m,C,sampler', [(26, 30, 40, 'NUTS'), (12, 17, 21, 'NUTS'), (16, 9, 16, 'HMC')])
def test_likelihood_tensor_core_samplers(handle, oracle, monkeypatch, n, m, C, sampler):
    """NUTS / HMC of the likelihood pipeline on the tensor-core kernels (model variant bit 3 of bfb_dmma.cuh: the outputs'
    operand streamed from L2): integer outcomes identical to the oracle fed with the device's own draws and to the generic
    kernel"""
    from bayesfast_b200.density import whiten_spec, GaussianLikelihood
    rng = np.random.default_rng(300 + n + m)
    spec = _lik_spec(rng, n, m)
    ep = spec['epilogue']
    handle.set_model(whiten_spec(to_device_spec(spec), GaussianLikelihood(ep['d'], ep['cinv'], ep['c0'])))
    n_iter, seed = 30, 909
    kw = {'n_int_step': 8} if sampler == 'HMC' else {}
    x0 = rng.normal(size=(C, n)) * 0.3
    step0 = 0.3 if sampler == 'HMC' else 1. / n**0.25
    outs = {}
    for fam in ('dmma', 'generic'):
        monkeypatch.setenv('BFB200_SAMPLER', fam)
        handle.sampler_init(cfg_from(kw, n_iter // 2, seed), x0, step0, np.ones(n), x0)
        outs[fam] = handle.sampler_run(sampler, n_iter)
        assert handle.sampler_last_path() == fam
        st = handle.sampler_state()
        assert np.all(st['status'] == 0)
    monkeypatch.delenv('BFB200_SAMPLER')
    U, Z = device_draws(handle, seed, st['n_draws'])
    ref = oracle.OracleDensity(spec).run(sampler, dict(n_iter=n_iter, n_warmup=n_iter // 2, **kw), x0, step0, np.ones(n),
                                         draws_u=U, draws_z=Z)
    for fam in ('dmma', 'generic'):
        for k in INT_STATS:
            assert np.array_equal(outs[fam][k], ref[k]), (fam, k)
        check_floats(outs[fam]['samples'], ref['samples'], fam)


DES = gio.load('pipeline_des.npz')['cases'][0]


@pytest.mark.parametrize('path', ['tensor', 'tensor_dense', 'generic'])
def test_pipeline_des_shaped_golden(handle, monkeypatch, path):
    """The DES-Y1 example's three-module density (examples/des-y1-w-cosmosis.ipynb cells 12-18: linear + shared-mask quadratic
    surrogate with module input_scales -> chi2 -> posterior module with a Gaussian prior on 13 inputs; Density input_scales,
    hard_bounds=True, radial bound) against the real reference: logp_and_grad at 1e-10 (6 of the 14 points lie outside the
    bound) and a NUTS run with identical integer outcomes -- on the tensor-core kernels (model variant bits 3 | 1) and on the
    generic ones."""
    if path == 'generic':
        monkeypatch.setenv('BFB200_EVAL', 'generic')
        monkeypatch.setenv('BFB200_SAMPLER', 'generic')
    if path == 'tensor_dense':
        monkeypatch.setenv('BFB200_LIK_DENSE', '1')       # one n x n product per output instead of the feature form Phi(x) C^T
    case = DES
    handle.set_model(_whitened(case))
    lp, g = handle.logp_and_grad_batch(case['X'])
    assert handle.eval_last_path() == {'generic': 'generic', 'tensor': 'lik_feat', 'tensor_dense': 'lik_dmma'}[path]
    assert np.allclose(lp, case['logp'], rtol=1e-10, atol=1e-10)
    assert np.allclose(g, case['grad'], rtol=1e-10, atol=1e-10 * np.abs(case['grad']).max())
    r, kw = case['result'], case['trace_kw']
    n_iter, n_warmup = int(kw['n_iter']), int(kw['n_warmup'])
    handle.sampler_init(cfg_from(kw, n_warmup, int(case['seed'])), case['x0'], float(r['step0']), r['var0'], case['x0'])
    out = handle.sampler_run('NUTS', n_iter)
    assert handle.sampler_last_path() == ('generic' if path == 'generic' else 'dmma')
    st = handle.sampler_state()
    assert np.all(st['status'] == 0) and np.array_equal(st['n_draws'], r['n_draws'])
    for k in INT_STATS:
        assert np.array_equal(out[k], r[k].astype(np.int32)), k
    assert int(out['diverging'].sum()) == int(r['diverging'].sum()) >= 1
    for k in FLT_STATS:
        assert np.allclose(out[k][:, :4], r[k][:, :4], rtol=1e-9, atol=1e-9), k
    assert np.allclose(out['samples'][:, :4], r['samples'][:, :4], rtol=1e-9, atol=1e-9)
    assert np.allclose(out['samples'], r['samples'], rtol=1e-3, atol=1e-3)


def test_des_shaped_public_api_and_large_batch(oracle):
    """the same density shape through the public interface (PolyModel with a shared-mask quadratic config, GaussianLikelihood,
    GaussianPrior, input_scales + hard_bounds=True): fit on the device, batched logp_and_grad of 3000 points (ragged) on the
    tensor-core evaluator against the oracle evaluating the device-fitted coefficients, HMC and NUTS decisions against the
    oracle fed with the device's draws"""
    import bayesfast_b200 as bfb
    sp = DES['spec']
    n, m = int(sp['n']), int(sp['m'])
    rng = np.random.default_rng(4)
    rg = sp['transform_ranges']
    mid, width = rg.mean(axis=1), rg[:, 1] - rg[:, 0]
    nl = np.asarray(sp['configs'][1]['input_mask'])
    W1, W2 = rng.normal(size=(m, n)), rng.normal(size=(m, nl.size, nl.size)) * 0.5
    sur = bfb.PolyModel([bfb.PolyConfig('linear'), bfb.PolyConfig('quadratic', input_mask=nl)], input_size=n, output_size=m,
                        input_scales=rg)
    pr = sp['prior']
    d_vec = rng.normal(size=m) * 0.1
    den = bfb.Density(sur, input_scales=rg, hard_bounds=True, likelihood=bfb.GaussianLikelihood(d_vec, np.ones(m), 1.5),
                      prior=bfb.GaussianPrior(pr['idx'], pr['mu'], pr['sig'], float(pr['c0'])))
    xf = mid + np.clip(rng.normal(size=(3 * sur.n_param, n)) * 0.08, -0.45, 0.45) * width
    u = (xf - mid) / width
    yf = u @ W1.T + np.einsum('ojk,nj,nk->no', W2, u[:, nl], u[:, nl])
    den.fit(xf, yf)
    od = oracle.OracleDensity(den.to_spec())
    Xo = mid + np.clip(rng.normal(size=(3000 - 7, n)) * np.where(np.arange(3000 - 7) % 2, 0.05, 0.15)[:, None], -0.49, 0.49) * width
    Xt = den.from_original(Xo)
    lp, g = den.logp_and_grad(Xt, original_space=False)
    assert den._sync(False).eval_last_path() == 'lik_feat'
    lpo, go = od.logp_and_grad_batch(Xt)
    assert np.allclose(lp, lpo, rtol=1e-10, atol=1e-9) and np.allclose(g, go, rtol=1e-9, atol=1e-10 * np.abs(go).max())
    beta = np.sqrt(np.einsum('ij,jk,ik->i', (Xo - rg[:, 0]) / width - sur._mu, sur._hess, (Xo - rg[:, 0]) / width - sur._mu))
    assert 0.05 < np.mean(beta > sur._alpha) < 0.95                   # both sides of the radial bound
    x0 = den.from_original(xf[:24])
    for sampler, kw in (('NUTS', {}), ('HMC', dict(n_int_step=6))):
        tt = bfb.sample(den, dict(n_chain=24, n_iter=30, n_warmup=15, x_0=xf[:24], random_generator=9, **kw),
                        sampler=sampler, verbose=False)
        assert den._sync(False).sampler_last_path() == 'dmma'
        U, Z = device_draws(den._sync(False), 9, tt._final['n_draws'])
        ref = od.run(sampler, dict(n_iter=30, n_warmup=15, **kw), x0, 1. / n**0.25, np.ones(n), draws_u=U, draws_z=Z)
        assert np.array_equal(tt._final['n_draws'], ref['n_draws'])
        assert np.array_equal(tt.arrays['tree_depth'], ref['tree_depth']) and np.array_equal(tt.arrays['diverging'], ref['diverging'])
        assert np.allclose(tt.samples[:, :4], ref['samples'][:, :4], rtol=1e-8, atol=1e-8)
