"""
Synthetic posteriors of the benchmark configurations (BASELINE.json configs), pure numpy:
true log-density, training points for the surrogate fit and chain starting points, all from fixed seeds.
"""
import numpy as np

__all__ = ['des_shaped', 'correlated_gaussian', 'des_pipeline', 'des_y1_like', 'cubic3_stack', 'n_param']


def n_param(order, n):
    p = n + 1
    if order in ('quadratic', 'cubic-2', 'cubic-3'):
        p += n * (n + 1) // 2
    if order in ('cubic-2', 'cubic-3'):
        p += n * n
    if order == 'cubic-3':
        p += n * (n - 1) * (n - 2) // 6
    return p


def _spd(n, cond, rng):
    Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    ev = np.exp(np.linspace(0., np.log(cond), n))
    P = (Q * ev) @ Q.T
    return 0.5 * (P + P.T)


def des_shaped(n=26, seed=0, cond=30., n_fit_mult=4, n_chain=4096, order='cubic-2'):
    """config 3: logp = -1/2 x^T P x - 0.02 sum x^3 exp(-0.1 x^2), P random SPD (cond ~10-100)."""
    rng = np.random.default_rng(seed)
    P = _spd(n, cond, rng)
    cov = np.linalg.inv(P)
    L = np.linalg.cholesky(cov)

    def logp(x):
        x = np.atleast_2d(x)
        return -0.5 * np.einsum('ij,jk,ik->i', x, P, x) - 0.02 * np.sum(x**3 * np.exp(-0.1 * x**2), axis=1)

    N = n_fit_mult * n_param(order, n)
    x_fit = (L @ rng.normal(size=(n, N))).T
    x_0 = (L @ rng.normal(size=(n, n_chain))).T
    return dict(n=n, order=order, logp=logp, x_fit=x_fit, y_fit=logp(x_fit)[:, None], x_0=x_0, cov=cov, P=P)


def correlated_gaussian(n=16, seed=0, n_fit_mult=4, n_chain=4096, order='cubic-2'):
    """config 2: Sigma = A A^T / n + I, exact Gaussian logp."""
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n))
    cov = A @ A.T / n + np.eye(n)
    P = np.linalg.inv(cov)
    L = np.linalg.cholesky(cov)

    def logp(x):
        x = np.atleast_2d(x)
        return -0.5 * np.einsum('ij,jk,ik->i', x, P, x)

    N = n_fit_mult * n_param(order, n)
    x_fit = (L @ rng.normal(size=(n, N))).T
    x_0 = (L @ rng.normal(size=(n, n_chain))).T
    return dict(n=n, order=order, logp=logp, x_fit=x_fit, y_fit=logp(x_fit)[:, None], x_0=x_0, cov=cov, P=P)


def des_pipeline(n=26, m=457, seed=0, n_blocks=8, n_in=10):
    """
    DES-Y1-shaped two-module pipeline (examples/des-y1-w-cosmosis.ipynb cells 12-18): m surrogate outputs in `n_blocks` blocks,
    each a quadratic in `n_in` of the n inputs (masked configs) on top of a full linear config, followed by a Gaussian
    likelihood with a dense inverse covariance.  Returns (spec, GaussianLikelihood): the spec carries dense `coef` (oracle),
    `packed` (C ABI) and the un-whitened `epilogue`.
    """
    from .density import GaussianLikelihood
    from .poly import pack_dense
    rng = np.random.default_rng(seed)
    edges = np.linspace(0, m, n_blocks + 1).astype(int)
    lin = np.concatenate((rng.normal(size=(m, 1)) * 0.1, rng.normal(size=(m, n)) * 0.3), axis=1)
    cfgs = [dict(order='linear', input_mask=np.arange(n), output_mask=np.arange(m), coef=lin, packed=lin.copy())]
    for b in range(n_blocks):
        k = edges[b + 1] - edges[b]
        if k == 0:
            continue
        ni = min(n_in, n)
        q = np.triu(rng.normal(size=(k, ni, ni))) * 0.03
        cfgs.append(dict(order='quadratic', input_mask=np.sort(rng.choice(n, size=ni, replace=False)),
                         output_mask=np.arange(edges[b], edges[b + 1]), coef=q,
                         packed=np.array([pack_dense('quadratic', qi, ni) for qi in q])))
    B = rng.normal(size=(m, m))
    lik = GaussianLikelihood(rng.normal(size=m) * 0.1, (B @ B.T / m + 0.5 * np.eye(m)) * (float(n) / m), 0.)
    spec = dict(n=n, m=m, configs=cfgs, use_bound=False, input_scales=None, use_decay=False, transform_ranges=None,
                epilogue=lik.to_spec())
    return spec, lik


def cubic3_stack(n=64, seed=3, cond=30., cubic_scale=0.02):
    """BASELINE configs[3]: a 64-D cubic-3 surrogate (linear + quadratic + cubic-2 + cubic-3 configs, P = 47905 at n = 64) with
    INJECTED coefficients -- a random SPD quadratic part (condition number `cond`) plus small random cubic terms, so that the
    curvature varies along a trajectory and the NUTS tree depths of the chains spread.  Returns (device spec with packed
    coefficients, covariance of the Gaussian part)."""
    rng = np.random.default_rng(seed)
    P = _spd(n, cond, rng)
    cov = np.linalg.inv(P)
    lin = np.zeros((1, n + 1))
    lin[0, 1:] = 0.05 * rng.normal(size=n)
    iu = np.triu_indices(n)
    quad = np.where(iu[0] == iu[1], -0.5 * P[iu], -P[iu])[None]
    c2 = (cubic_scale * rng.normal(size=(1, n * n)) / n)
    c3 = (cubic_scale * rng.normal(size=(1, n * (n - 1) * (n - 2) // 6)) / n)
    full = np.arange(n)
    cfgs = [dict(order=o, input_mask=full, output_mask=np.arange(1), packed=a)
            for o, a in (('linear', lin), ('quadratic', quad), ('cubic-2', c2), ('cubic-3', c3))]
    spec = dict(n=n, m=1, configs=cfgs, use_bound=False, input_scales=None, use_decay=False, transform_ranges=None)
    return spec, cov


DES_PARA_RANGE = np.array([[0.1, 0.9], [0.55, 0.9], [0.03, 0.07], [0.87, 1.07], [0.5e-9, 5.0e-9], [0.0006, 0.01], [-2, -0.333],
                           [0.8, 3.0], [0.8, 3.0], [0.8, 3.0], [0.8, 3.0], [0.8, 3.0], [-0.1, 0.1], [-0.1, 0.1], [-0.1, 0.1],
                           [-0.1, 0.1], [-5.0, 5.0], [-5.0, 5.0], [-0.1, 0.1], [-0.1, 0.1], [-0.1, 0.1], [-0.1, 0.1],
                           [-0.05, 0.05], [-0.05, 0.05], [-0.05, 0.05], [-0.05, 0.05], [-0.05, 0.05]])
DES_NONLINEAR = np.array([0, 1, 2, 3, 4, 5, 6, 16, 17])
DES_PRIOR = dict(idx=np.array([18, 19, 20, 21, 12, 13, 14, 15, 22, 23, 24, 25, 26]),
                 mu=np.array([-0.001, -0.019, 0.009, -0.018, 0.012, 0.012, 0.012, 0.012, 0.008, -0.005, 0.006, 0.0, 0.0]),
                 sig=np.array([0.016, 0.013, 0.011, 0.022, 0.023, 0.023, 0.023, 0.023, 0.007, 0.007, 0.006, 0.01, 0.01]))


def des_y1_like(m=457, seed=0, n_fit_mult=2):
    """The density shape of examples/des-y1-w-cosmosis.ipynb (cells 9-18) with a synthetic theory vector: n = 27 inputs with the
    notebook's parameter ranges (module rescale AND hard-bounded variable transform), m whitened outputs from a linear config
    plus a quadratic config on the notebook's shared 9-D mask, chi^2 likelihood, Gaussian prior on 13 inputs.
    Returns dict(ranges, nonlinear, prior, d, x_fit (original space), y_fit, x_0 (original space))."""
    rng = np.random.default_rng(seed)
    rg = DES_PARA_RANGE
    n = rg.shape[0]
    mid, width = rg.mean(axis=1), rg[:, 1] - rg[:, 0]
    W1 = rng.normal(size=(m, n)) * 1.5
    W2 = rng.normal(size=(m, 9, 9)) * 0.8

    def theory(x):
        u = (np.atleast_2d(x) - mid) / width
        v = u[:, DES_NONLINEAR]
        return u @ W1.T + np.einsum('ojk,nj,nk->no', W2, v, v)

    P = n + 1 + 9 * 10 // 2
    x_fit = mid + np.clip(rng.normal(size=(n_fit_mult * P, n)) * 0.08, -0.45, 0.45) * width
    d = theory(mid + rng.normal(size=n) * 0.03 * width)[0] + 0.05 * rng.normal(size=m)
    x_0 = mid + np.clip(rng.normal(size=(4096, n)) * 0.03, -0.45, 0.45) * width
    return dict(n=n, m=m, ranges=rg, nonlinear=DES_NONLINEAR, prior=DES_PRIOR, d=d, theory=theory, x_fit=x_fit, y_fit=theory(x_fit),
                x_0=x_0)
