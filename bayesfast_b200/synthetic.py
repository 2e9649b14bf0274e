"""
Synthetic posteriors of the benchmark configurations (BASELINE.json configs), pure numpy:
true log-density, training points for the surrogate fit and chain starting points, all from fixed seeds.
"""
import numpy as np

__all__ = ['des_shaped', 'correlated_gaussian', 'n_param']


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
