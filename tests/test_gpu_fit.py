"""GPU parity of PolyModel.fit (fused feature expansion + DMMA Gram + Cholesky) vs the reference's lstsq results."""
import warnings

import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import pytest

import _golden_io as gio

pytestmark = pytest.mark.gpu
# BASELINE.json north_star: surrogate coefficients and values within 1e-10 relative error in FP64.
# Coefficients are compared norm-wise per config block (|dc|_inf / |c|_inf); values point-wise.
COEF_TOL = 1e-10
VAL_TOL = 1e-10


def build_model(spec, **kw):
    import bayesfast_b200 as bfb
    cfgs = [bfb.PolyConfig(c['order'], c['input_mask'], c['output_mask']) for c in spec['configs']]
    return bfb.PolyModel(cfgs, input_size=int(spec['n']), output_size=int(spec['m']), **kw)


def block_err(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


def test_fit_golden(oracle):
    for c in gio.load('fit.npz')['cases']:
        spec = c['spec']
        s = build_model(spec, bound_options=dict(alpha_p=float(c['alpha_p']), center_max=bool(c['center_max'])))
        s.fit(c['x'], c['y'], c['logp'], c['w'])
        for conf, cf in zip(s.configs, spec['configs']):
            assert block_err(conf._coef, cf['coef']) < COEF_TOL, (c['name'], conf.order, block_err(conf._coef, cf['coef']))
        assert np.allclose(s._mu, spec['mu'], rtol=1e-12, atol=1e-14), c['name']
        assert block_err(s._hess, spec['hess']) < 1e-11, c['name']
        assert abs(s._alpha - spec['alpha']) < 1e-11 * spec['alpha'], c['name']
        assert block_err(s._f_mu, spec['f_mu']) < 1e-9, c['name']
        # predictions of the fitted model vs the reference's fitted model at the training points
        F, _ = s.eval_batch(c['x'][:64])
        Fo, _ = oracle.OracleDensity(spec).poly_eval_batch(c['x'][:64])
        assert np.max(np.abs(F - Fo) / np.maximum(np.abs(Fo), 1e-3 * np.max(np.abs(Fo)))) < VAL_TOL, c['name']


def test_reference_test_poly_case():
    """bayesfast/tests/test_poly.py:18-26 with this package in place of bayesfast"""
    import bayesfast_b200 as bfb
    g = gio.load('poly_kat.npz')
    x, y = g['x'], g['y']
    s = bfb.PolyModel('cubic-3', input_size=4, output_size=1)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        s.fit(x, y)
    y_s = np.concatenate([s(x_i)[0] for x_i in x])
    assert np.isclose(y_s, y[:, 0]).all()
    assert np.max(np.abs(y_s - y[:, 0])) < 1e-11
    s.fit(x, y, logp=y[:, 0])
    assert abs(s._alpha - 3.557097890322362) < 1e-11
    assert abs(s._f_mu[0] - 13.784528300586311) < 1e-9
    j = s.jac(x[0])[0]
    assert np.allclose(j[0], [-2.5378094262371858, 7.096831063745606, -1.333398061939275, -1.894373524859588], rtol=1e-10)
    # exact generating polynomial recovered
    lin, quad, c2, c3 = [c._coef[0] for c in s.configs]
    assert abs(lin[0] + 8) < 1e-11 and abs(lin[2] - 7) < 1e-11 and abs(quad[0, 0] - 5) < 1e-11
    assert abs(quad[0, 2] + 6) < 1e-11 and abs(c2[0, 0] - 1) < 1e-11 and abs(c2[1, 1] + 2) < 1e-11
    assert abs(c2[2, 3] + 4) < 1e-11 and abs(c3[1, 2, 3] - 3) < 1e-11


@pytest.mark.parametrize('n,N,scale', [(26, 4216, 1.0), (16, 1636, 1.0), (26, 4216, 0.3), (32, 20000, 1.0)])
def test_fit_vs_lstsq(oracle, n, N, scale):
    """BASELINE config shapes: cubic-2, N = 4 P points; compared with the oracle's scipy.linalg.lstsq fit"""
    import bayesfast_b200 as bfb
    rng = np.random.default_rng(n)
    A = rng.normal(size=(n, n))
    P = A @ A.T / n + np.eye(n)
    x = rng.normal(size=(N, n)) * scale
    y = (-0.5 * np.einsum('ij,jk,ik->i', x, P, x) - 0.02 * np.sum(x**3 * np.exp(-0.1 * x**2), axis=1))[:, None]
    s = bfb.PolyModel('cubic-2', input_size=n, output_size=1)
    s.fit(x, y, logp=y[:, 0])
    cfgs = [dict(order=c.order, input_mask=c.input_mask, output_mask=c.output_mask) for c in s.configs]
    ref = oracle.fit(cfgs, n, 1, x, y)
    errs = [block_err(conf._coef, r) for conf, r in zip(s.configs, ref)]
    F, _ = s.eval_batch(x[:512])
    spec = s.to_spec(with_bound=False)
    for cf, r in zip(spec['configs'], ref):
        cf['coef'] = r
    Fo, _ = oracle.OracleDensity(spec).poly_eval_batch(x[:512])
    verr = np.max(np.abs(F - Fo) / np.maximum(np.abs(Fo), 1e-3 * np.max(np.abs(Fo))))
    print('n', n, 'N', N, 'scale', scale, 'coef errs', errs, 'value err', verr, 'resid', s._fit_rel_resid)
    assert max(errs) < COEF_TOL          # measured 4e-12 .. 2e-11, also for the ill-scaled inputs (scale 0.3)
    assert verr < VAL_TOL
    mu, hess, alpha = oracle.bound_from_points(x)
    assert block_err(s._hess, hess) < 1e-10 and abs(s._alpha - alpha) < 1e-11 * alpha


def test_fit_errors():
    import bayesfast_b200 as bfb
    from bayesfast_b200 import _cabi
    s = bfb.PolyModel('cubic-2', input_size=4, output_size=1)
    x = np.random.default_rng(0).normal(size=(10, 4))
    with pytest.raises(ValueError):
        s.fit(x, x[:, :1])                     # fewer points than parameters (poly.py:521-523)
    with pytest.raises(ValueError):
        s.fit(x, x[:, :2])
    xd = np.repeat(x[:3], 20, axis=0)          # rank deficient
    with pytest.raises(_cabi.BfbError):
        s.fit(xd, xd[:, :1])


def test_density_fit_and_sample():
    """Recipe-shaped flow: evaluate the true logp on the host, fit the surrogate on the GPU, sample on the GPU."""
    import bayesfast_b200 as bfb
    n = 8
    rng = np.random.default_rng(1)
    A = rng.normal(size=(n, n))
    cov = A @ A.T / n + np.eye(n)
    Pm = np.linalg.inv(cov)
    xf = (np.linalg.cholesky(cov) @ rng.normal(size=(n, 2000))).T
    logp = -0.5 * np.einsum('ij,jk,ik->i', xf, Pm, xf)
    den = bfb.Density(bfb.PolyModel('quadratic', input_size=n, output_size=1), decay_options=dict(use_decay=True))
    den.fit(xf, logp)
    lp, g = den.logp_and_grad(xf[:5], original_space=False)
    assert np.allclose(lp, logp[:5], rtol=1e-9, atol=1e-9) and np.allclose(g, -xf[:5] @ Pm, rtol=1e-8, atol=1e-8)
    tt = bfb.sample(den, dict(n_chain=256, n_iter=400, n_warmup=200, x_0=xf[:256], random_generator=3), verbose=False)
    post = tt.get()
    assert post.shape == (256 * 200, n)
    assert np.all(np.abs(post.mean(axis=0)) < 0.05 * np.sqrt(np.diag(cov)) * 3)
    assert np.allclose(np.cov(post, rowvar=False), cov, atol=0.08 * np.max(np.diag(cov)))


def _sharded_fit_worker(rank, world, port, q):
    import os, sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)        # two ranks share cuda:0: the exchange is host-staged
    import bayesfast_b200 as bfb
    from bayesfast_b200.runtime import shard_bounds
    x, y = _sharded_problem()
    lo, hi = shard_bounds(x.shape[0], rank, world) if rank >= 0 else (0, x.shape[0])
    if rank == 1:
        lo, hi = lo + 37, hi          # uneven shards
    elif rank == 0:
        hi = hi + 37
    sur = bfb.PolyModel('cubic-2', input_size=x.shape[1], output_size=1)
    den = bfb.Density(sur, decay_options=dict(use_decay=True))
    den.fit(x[lo:hi], y[lo:hi], comm=True)
    q.put((rank, [np.array(c._coef) for c in sur.configs], sur._mu, sur._hess, sur._alpha, sur._f_mu, den._mu, den._hess,
           den._alpha, dict(sur._fit_exchange)))
    dist.barrier()
    dist.destroy_process_group()


def _sharded_problem():
    rng = np.random.default_rng(21)
    n, N = 7, 900
    x = rng.normal(size=(N, n)) * 0.8 + 0.1
    y = (-0.5 * np.sum(x**2, axis=1) + 0.1 * x[:, 0] * x[:, 1]**2 + 0.01 * rng.normal(size=N))[:, None]
    return x, y


def test_sharded_fit_product_path_two_ranks():
    """Density.fit(..., comm=True) of two ranks (gloo, both on cuda:0) against the single-process fit: the packed exchange
    buffer (upper block triangle of the Gram + moments), the all-reduced bound radius, the center_max vote and the decay
    ellipsoid give every rank the SAME model as one process with all rows"""
    import torch.multiprocessing as mp
    import bayesfast_b200 as bfb
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sharded_fit_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y = _sharded_problem()
    sur = bfb.PolyModel('cubic-2', input_size=x.shape[1], output_size=1)
    den = bfb.Density(sur, decay_options=dict(use_decay=True))
    den.fit(x, y)
    for r in res:
        for a, b in zip(r[1], [np.array(c._coef) for c in sur.configs]):
            assert np.allclose(a, b, rtol=1e-9, atol=1e-11 * np.abs(b).max())
        assert np.allclose(r[2], sur._mu, rtol=1e-12, atol=1e-14) and np.allclose(r[3], sur._hess, rtol=1e-10)
        assert np.isclose(r[4], sur._alpha, rtol=1e-10) and np.allclose(r[5], sur._f_mu, rtol=1e-9)
        assert np.allclose(r[6], den._mu, rtol=1e-12, atol=1e-14) and np.allclose(r[7], den._hess, rtol=1e-10)
        assert np.isclose(r[8], den._alpha, rtol=1e-10)
        assert r[9]['allreduce_bytes'] < 8 * (2 * 64 * 64 * 3 // 2 + 64)       # packed upper triangle, not the full P x P
    for k in range(1, 9):                                # identical on both ranks: no broadcast needed after the solve
        a, b = res[0][k], res[1][k]
        assert all(np.array_equal(u, v) for u, v in zip(a, b)) if isinstance(a, list) else np.array_equal(a, b)


def test_post_step_on_device(oracle):
    """importance weights + truncation and SystematicResampler on the device (bitonic argsort, csrc/bfb_post.cu)
    against the real reference's outputs (tests/golden/post.npz), against the oracle at sizes with ragged / non-power-of-two
    lengths, and with device-resident (torch) inputs"""
    import torch
    from bayesfast_b200.post import SystematicResampler, importance_weights
    g = gio.load('post.npz')
    for c in g['cases']:
        rs = SystematicResampler(nodes=c['nodes'], weights=c['weights'])
        assert np.array_equal(rs.run(c['a'], int(c['n'])), c['idx']), c['name']
        d = rs.run(torch.from_numpy(c['a']).cuda(), int(c['n']))
        assert d.is_cuda and np.array_equal(d.cpu().numpy(), c['idx']), c['name']
    w, wt, st = importance_weights(g['logp'], g['logq'], float(g['k_trunc']))
    assert np.allclose(w, g['weights'], rtol=1e-14) and np.allclose(wt, g['weights_trunc'], rtol=1e-13)
    assert np.isclose(st['sum'], g['weights'].sum(), rtol=1e-13) and np.isclose(st['max_trunc'], g['weights_trunc'].max(), rtol=1e-13)
    rng = np.random.default_rng(9)
    for m, n in ((1, 1), (2047, 100), (2049, 333), (1 << 20, 5000), (3000001, 47905)):
        a = rng.normal(size=m)
        a[rng.integers(0, m, size=min(m, 50))] = a[0]                       # ties: ordered by index like a stable argsort
        rs = SystematicResampler(require_unique=False)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            got = rs.run(torch.from_numpy(a).cuda(), n).cpu().numpy()
        ref, _ = oracle.systematic_resample(a, n)
        assert np.array_equal(got, ref), (m, n)
    lp, lq = torch.randn(1 << 21, dtype=torch.float64, device='cuda'), torch.randn(1 << 21, dtype=torch.float64, device='cuda')
    for kt in (0.25, -1.):
        w, wt, st = importance_weights(lp, lq, kt)
        wo, wto = oracle.importance_weights(lp.cpu().numpy(), lq.cpu().numpy(), kt)
        assert np.allclose(w.cpu().numpy(), wo, rtol=1e-14) and np.allclose(wt.cpu().numpy(), wto, rtol=1e-12)
        assert np.isclose(st['n_eff'], wto.sum()**2 / (wto**2).sum(), rtol=1e-10)
    with pytest.raises(RuntimeError):
        SystematicResampler().run(np.arange(10.), 50)                        # not unique (utils/misc.py:102-106)
    with pytest.raises(ValueError):
        SystematicResampler(nodes=(5., 1.))


def test_fit_from_device_resident_rows():
    """PolyModel.fit on CUDA torch tensors (rows that never left the GPU) == the same fit from host arrays, bound included"""
    import torch
    import bayesfast_b200 as bfb
    rng = np.random.default_rng(31)
    n, N = 9, 1500
    x = rng.normal(size=(N, n)) * 0.7 + 0.2
    y = (-0.5 * np.sum(x**2, axis=1) + 0.2 * x[:, 1] * x[:, 2]**2 + 0.01 * rng.normal(size=N))[:, None]
    w = rng.uniform(0.5, 1.5, size=N)
    a = bfb.PolyModel('cubic-2', input_size=n, output_size=1)
    a.fit(x, y, logp=y[:, 0], w=w)
    b = bfb.PolyModel('cubic-2', input_size=n, output_size=1)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    b.fit(xd, yd, logp=yd[:, 0].contiguous(), w=torch.from_numpy(w).cuda())
    for ca, cb in zip(a.configs, b.configs):
        assert np.array_equal(np.asarray(ca._coef), np.asarray(cb._coef))
    assert np.array_equal(a._mu, b._mu) and np.array_equal(a._hess, b._hess) and a._alpha == b._alpha
    assert np.array_equal(a._f_mu, b._f_mu)
