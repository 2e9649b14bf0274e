"""Spec helpers shared by the tests: golden/oracle specs carry dense `coef`; the C ABI wants `packed`."""
import numpy as np


def pack(order, c, n):
    c = np.asarray(c, dtype=np.float64)
    if order == 'linear':
        return c.copy()
    if order == 'quadratic':
        return c[np.triu_indices(n)].copy()
    if order == 'cubic-2':
        return c.reshape(n * n).copy()
    out = []
    for j in range(n):
        for k in range(j + 1, n):
            out.append(c[j, k, k + 1:])
    return np.concatenate(out) if out else np.zeros(0)


def to_device_spec(spec):
    s = dict(spec)
    s['configs'] = []
    for cf in spec['configs']:
        d = dict(cf)
        ni = len(cf['input_mask'])
        d['packed'] = np.array([pack(cf['order'], ci, ni) for ci in np.asarray(cf['coef'])])
        s['configs'].append(d)
    return s


def gpu_available():
    try:
        from bayesfast_b200 import _cabi
        return _cabi.lib().bfb_device_count() > 0
    except Exception:
        return False


def synthetic_spec(n, order='cubic-2', seed=0, cubic_scale=0.02, cond=30., bound=True, decay=False, transform=False):
    """
    A well-behaved synthetic posterior in surrogate form (SURVEY.md 8d, config 3 shape):
    logp = -1/2 x^T P x + small cubic terms, P random SPD with the given condition number.
    Returns (spec with dense `coef`, covariance of the Gaussian part).
    """
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    ev = np.exp(np.linspace(0., np.log(cond), n))
    P = (Q * ev) @ Q.T
    P = 0.5 * (P + P.T)
    cov = np.linalg.inv(P)
    lin = np.zeros((1, n + 1))
    lin[0, 1:] = 0.05 * rng.normal(size=n)
    quad = np.zeros((1, n, n))
    for j in range(n):
        quad[0, j, j] = -0.5 * P[j, j]
        quad[0, j, j + 1:] = -P[j, j + 1:]
    cfgs = [dict(order='linear', input_mask=np.arange(n), output_mask=np.arange(1), coef=lin),
            dict(order='quadratic', input_mask=np.arange(n), output_mask=np.arange(1), coef=quad)]
    if order in ('cubic-2', 'cubic-3'):
        c2 = cubic_scale * rng.normal(size=(1, n, n)) / n
        cfgs.append(dict(order='cubic-2', input_mask=np.arange(n), output_mask=np.arange(1), coef=c2))
    if order == 'cubic-3':
        c3 = np.zeros((1, n, n, n))
        for j in range(n):
            for k in range(j + 1, n):
                c3[0, j, k, k + 1:] = cubic_scale * rng.normal(size=n - k - 1) / n
        cfgs.append(dict(order='cubic-3', input_mask=np.arange(n), output_mask=np.arange(1), coef=c3))
    spec = dict(n=n, m=1, configs=cfgs, use_bound=bool(bound), input_scales=None, use_decay=bool(decay),
                transform_ranges=None)
    L = np.linalg.cholesky(cov)
    pts = (L @ rng.normal(size=(n, 4 * n + 50))).T
    mu = pts.mean(axis=0)
    hess = np.linalg.inv(np.cov(pts, rowvar=False))
    beta = np.einsum('ij,jk,ik->i', pts - mu, hess, pts - mu) ** 0.5
    if bound:
        spec.update(mu=mu, hess=hess, alpha=float(beta.max()) * 1.6, f_mu=np.zeros(1))
    if decay:
        spec.update(d_mu=mu, d_hess=hess, d_alpha2=float(beta.max() * 1.5) ** 2, d_gamma=0.1)
    if transform:
        sd = np.sqrt(np.diag(cov))
        ranges = np.stack((-12. * sd, 12. * sd), axis=1)
        hb = np.zeros((n, 2), np.uint8)
        hb[0] = (1, 1)
        hb[min(2, n - 1)] = (1, 0)
        spec.update(transform_ranges=ranges, hard_bounds=hb)
    return spec, cov


def unpack(order, a, n):
    """inverse of pack(): dense reference-layout tensor from the packed coefficient vector (PolyConfig._set, poly.py:131-158)"""
    a = np.asarray(a, dtype=np.float64)
    if order == 'linear':
        return a.copy()
    if order == 'quadratic':
        c = np.zeros((n, n))
        c[np.triu_indices(n)] = a
        return c
    if order == 'cubic-2':
        return a.reshape(n, n).copy()
    c = np.zeros((n, n, n))
    i = 0
    for j in range(n):
        for k in range(j + 1, n):
            m = n - k - 1
            c[j, k, k + 1:] = a[i:i + m]
            i += m
    return c


def c3n64_spec(g):
    """the 64-D cubic-3 stack of tests/golden/poly_eval_c3n64.npz: coefficients regenerated from the recorded seed (the same
    numpy calls as make_golden.seeded_coefs), bound parameters from the fixture"""
    n = int(g['n'])
    orders = ('linear', 'quadratic', 'cubic-2', 'cubic-3')
    shapes = {'linear': n + 1, 'quadratic': n * (n + 1) // 2, 'cubic-2': n * n, 'cubic-3': n * (n - 1) * (n - 2) // 6}
    rng = np.random.default_rng(int(g['seed']))
    cfgs = []
    for o in orders:
        a = rng.normal(size=shapes[o]) * float(g['scales'][o])
        cfgs.append(dict(order=o, input_mask=np.arange(n), output_mask=np.arange(1), coef=unpack(o, a, n)[None]))
    return dict(n=n, m=1, configs=cfgs, use_bound=True, mu=g['mu'], hess=g['hess'], alpha=float(g['alpha']), f_mu=g['f_mu'],
                input_scales=None, use_decay=False, transform_ranges=None)
