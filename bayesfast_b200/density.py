"""
Surrogate-only probability density with the interface of bayesfast.core.density.Density
(reference: bayesfast/core/density.py:617-838 on top of Pipeline :205-566 and _PipelineBase :28-157).

The reference's Density chains arbitrary Python modules; only a density whose logp is produced by ONE
PolyModel surrogate -- directly (output #0), or through one closed-form Gaussian-likelihood module placed after it
(`GaussianLikelihood`: f_1 of examples/2d-donut.ipynb, the chi-square module of examples/des-y1-w-cosmosis.ipynb) -- can run
on the device, so this class holds exactly that: the surrogate, the optional likelihood,
the optional bounded<->unbounded variable transform (input_scales / hard_bounds) and the optional decay
term.  `Density.from_reference(d)` builds one from a fitted reference Density (duck-typed, no import of the
reference) and refuses anything else.
"""
from collections import namedtuple

import numpy as np

from . import transforms as tf
from ._cabi import n_packed
from .poly import PolyModel, PolyConfig, pack_dense, unpack_dense

__all__ = ['Density', 'DecayOptions', 'GaussianLikelihood', 'GaussianPrior', 'whiten_spec']

DecayOptions = namedtuple('DecayOptions', ('use_decay', 'alpha', 'alpha_p', 'gamma'))


class GaussianLikelihood:
    """
    The second module of a two-module pipeline, m surrogate outputs -> logp:

        logp(f) = const - 1/2 (f - mean)^T inv_cov (f - mean)

    In the reference this is a user `bayesfast.Module(fun=..., jac=...)` placed after the surrogate in
    `Density(module_list=[...])` (core/density.py:487-566 chains the Jacobians): f_1 = -(m - 5)^2 / 0.5 of
    examples/2d-donut.ipynb is GaussianLikelihood([5.], [[4.]]), the chi^2 + prior module of
    examples/des-y1-w-cosmosis.ipynb is one with the data vector and the inverse covariance.  Arbitrary Python
    modules cannot run on the device; this closed form can.
    """

    def __init__(self, mean, inv_cov, const=0.):
        self.mean = np.atleast_1d(np.asarray(mean, dtype=np.float64))
        m = self.mean.size
        ic = np.asarray(inv_cov, dtype=np.float64)
        if ic.ndim == 1:
            ic = np.diag(ic)
        ic = np.atleast_2d(ic)
        if self.mean.ndim != 1 or ic.shape != (m, m):
            raise ValueError('mean should have shape (m,) and inv_cov (m, m) or (m,).')
        self.inv_cov = ic.copy()
        self.const = float(const)
        lam, V = np.linalg.eigh(0.5 * (ic + ic.T))
        if not np.all(lam >= -1e-12 * max(1., np.abs(lam).max())):
            raise ValueError('inv_cov should be positive semi-definite.')
        # symmetric part of inv_cov = Lt^T Lt (rows belonging to zero eigenvalues are zero and contribute nothing)
        self._lt = np.sqrt(np.clip(lam, 0., None))[:, None] * V.T

    output_size = property(lambda self: self.mean.size)

    def logp(self, f):
        r = np.asarray(f, dtype=np.float64) - self.mean
        return self.const - 0.5 * np.einsum('...i,ij,...j->...', r, self.inv_cov, r)

    __call__ = logp

    def to_spec(self):
        return dict(d=self.mean.copy(), cinv=self.inv_cov.copy(), c0=self.const)


class GaussianPrior:
    """
    The third module of the DES-Y1 example's pipeline (examples/des-y1-w-cosmosis.ipynb cells 12-14): an independent Gaussian
    prior on some of the ORIGINAL-space inputs added to the likelihood,

        logp = like + const - 1/2 sum_k ((x[indices[k]] - mean[k]) / sigma[k])^2

    (des_post_f / des_post_fj, a user Module with inputs ['like', 'x'] in the reference).  It is evaluated at the point itself,
    not at its projection onto the surrogate's radial bound, and its Jacobian is chained with the variable transform.
    """

    def __init__(self, indices, mean, sigma, const=0.):
        self.indices = np.atleast_1d(np.asarray(indices, dtype=np.int64))
        self.mean = np.atleast_1d(np.asarray(mean, dtype=np.float64))
        self.sigma = np.atleast_1d(np.asarray(sigma, dtype=np.float64))
        if not (self.indices.ndim == 1 and self.mean.shape == self.indices.shape == self.sigma.shape):
            raise ValueError('indices, mean and sigma should be 1-d arrays of the same length.')
        if np.unique(self.indices).size != self.indices.size or np.any(self.indices < 0):
            raise ValueError('indices should be unique and non-negative.')
        if not np.all(self.sigma > 0):
            raise ValueError('sigma should be positive.')
        self.const = float(const)

    def logp(self, x):
        x = np.asarray(x, dtype=np.float64)
        return self.const - 0.5 * np.sum(((x[..., self.indices] - self.mean) / self.sigma)**2, axis=-1)

    __call__ = logp

    def to_spec(self):
        return dict(idx=self.indices.copy(), mu=self.mean.copy(), sig=self.sigma.copy(), c0=self.const)


def whiten_spec(spec, lik):
    """
    Fold a GaussianLikelihood into the surrogate: f'(x) = Lt (f(x) - mean) with inv_cov = Lt^T Lt is again a polynomial
    of the same orders (every coefficient, the constants and the bound's f_mu are linear in the outputs), so that
    logp = const - 1/2 sum_o f'_o(x)^2 and grad = -sum_o f'_o grad f'_o: the device then needs no m x m product.
    Returns a spec with one full-input config per order present (all m outputs), `epilogue_sumsq` = const.
    """
    n, m = int(spec['n']), int(spec['m'])
    if lik.output_size != m:
        raise ValueError('the likelihood takes {} inputs but the surrogate has {} outputs.'.format(lik.output_size, m))
    full = {}
    for c in spec['configs']:
        order, im, om = c['order'], np.asarray(c['input_mask'], np.int64), np.asarray(c['output_mask'], np.int64)
        if c.get('packed', None) is None:
            raise RuntimeError('the surrogate has no coefficients yet.')
        pk = np.asarray(c['packed'], np.float64).reshape(om.size, n_packed(order, im.size))
        acc = full.setdefault(order, np.zeros((m, n_packed(order, n))))
        for i, o in enumerate(om):
            dense = unpack_dense(order, pk[i], im.size)
            if order == 'linear':
                big = np.zeros(n + 1)
                big[0] = dense[0]
                big[1 + im] = dense[1:]
            else:
                big = np.zeros((n,) * dense.ndim)
                big[np.ix_(*([im] * dense.ndim))] = dense
            acc[o] += pack_dense(order, big, n)
    lin = full.setdefault('linear', np.zeros((m, n + 1)))
    lin[:, 0] -= lik.mean
    lt = lik._lt
    out = dict(spec)
    out['configs'] = [dict(order=o, input_mask=np.arange(n), output_mask=np.arange(m), packed=lt @ full[o], coef=None)
                      for o in ('linear', 'quadratic', 'cubic-2', 'cubic-3') if o in full]
    if spec.get('use_bound', False):
        out['f_mu'] = lt @ (np.atleast_1d(spec['f_mu']).astype(np.float64) - lik.mean)
    out['epilogue_sumsq'] = lik.const
    out.pop('epilogue', None)
    return out


class Density:
    """
    Parameters
    ----------
    surrogate : PolyModel
        Its output #0 is the logarithmic density (density.py:737-739).
    input_scales, hard_bounds : as in the reference's Pipeline (density.py:225-233)
    decay_options : dict, optional
        Passed to set_decay_options (density.py:761-794).
    """

    def __init__(self, surrogate, input_scales=None, hard_bounds=False, decay_options=None,
                 density_name='__var__', input_vars='__var__', likelihood=None, prior=None):
        if not isinstance(surrogate, PolyModel):
            raise ValueError('surrogate should be a bayesfast_b200.PolyModel: only surrogate-only densities run '
                             'on the device, there is no CPU fallback.')
        self._surrogate = surrogate
        if likelihood is not None and not isinstance(likelihood, GaussianLikelihood):
            raise ValueError('likelihood should be a GaussianLikelihood: arbitrary Python modules cannot run on the device.')
        if likelihood is not None and likelihood.output_size != surrogate.output_size:
            raise ValueError('the likelihood takes {} inputs but the surrogate has {} outputs.'.format(
                likelihood.output_size, surrogate.output_size))
        self._likelihood = likelihood
        if prior is not None and not isinstance(prior, GaussianPrior):
            raise ValueError('prior should be a GaussianPrior: arbitrary Python modules cannot run on the device.')
        if prior is not None and prior.indices.size and prior.indices.max() >= surrogate.input_size:
            raise ValueError('the prior refers to input {} but the density has {} inputs.'.format(int(prior.indices.max()), surrogate.input_size))
        self._prior = prior
        self.density_name = str(density_name)
        self.input_vars = [input_vars] if isinstance(input_vars, str) else list(input_vars)
        n = surrogate.input_size
        self._input_scales = None if input_scales is None else tf.check_scales(input_scales)
        if self._input_scales is not None and self._input_scales.shape != (n, 2):
            raise ValueError('input_scales should have shape ({}, 2).'.format(n))
        self._hard_bounds = tf.check_bounds(hard_bounds, n)
        self.original_space = True
        self.use_surrogate = True
        self.set_decay_options(**({} if decay_options is None else decay_options))
        self._handle = None
        self._synced = set()
        self._dirty = True

    surrogate = property(lambda self: self._surrogate)
    surrogate_list = property(lambda self: [self._surrogate])
    likelihood = property(lambda self: self.__dict__.get('_likelihood', None))
    prior = property(lambda self: self.__dict__.get('_prior', None))
    input_size = property(lambda self: self._surrogate.input_size)
    input_scales = property(lambda self: self._input_scales)
    hard_bounds = property(lambda self: self._hard_bounds)

    # ------------------------------------------------------------------ decay (density.py:756-811)
    @property
    def decay_options(self):
        return DecayOptions(self._use_decay, self._alpha, self._alpha_p, self._gamma)

    def set_decay_options(self, use_decay=False, alpha=None, alpha_p=150., gamma=0.1):
        self._use_decay = bool(use_decay)
        if alpha is None:
            self._alpha = self._alpha_2 = None
        else:
            try:
                alpha = float(alpha)
                assert alpha > 0
            except Exception:
                raise ValueError('invalid value for alpha.')
            self._alpha, self._alpha_2 = alpha, alpha**2
        if alpha_p is None:
            if alpha is None:
                raise ValueError('alpha and alpha_p cannot both be None.')
            self._alpha_p = None
        else:
            try:
                alpha_p = float(alpha_p)
                assert alpha_p > 0
            except Exception:
                raise ValueError('invalid value for alpha_p.')
            self._alpha_p = alpha_p
        try:
            gamma = float(gamma)
            assert gamma > 0
        except Exception:
            raise ValueError('invalid value for gamma.')
        self._gamma = gamma
        self._dirty = True

    def _set_decay(self, x, comm=None):
        """ellipsoid of the fit points in the ORIGINAL space (density.py:796-811); with rows sharded over ranks (comm) the
        moments and the largest radius are all-reduced like PolyModel's bound, so every rank holds the same density"""
        from .fit import ellipsoid
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim != 2:
            raise ValueError('invalid value for x.')
        self._mu, self._hess, a = ellipsoid(self._surrogate, x, self._alpha_p, comm=comm)
        if self._alpha_p is not None:
            self._alpha, self._alpha_2 = a, a**2
        self._dirty = True

    # ------------------------------------------------------------------ transforms (density.py:92-203)
    def _tr(self, f, x, ident):
        x = np.asarray(x, dtype=np.float64)
        if self._input_scales is None:
            return ident(x)
        return f(x, self._input_scales, self._hard_bounds)

    def from_original(self, x):
        return self._tr(tf.from_original, x, np.copy)

    def to_original(self, x):
        return self._tr(tf.to_original, x, np.copy)

    def to_original_grad(self, x):
        return self._tr(tf.to_original_grad, x, np.ones_like)

    def to_original_grad2(self, x):
        return self._tr(tf.to_original_grad2, x, np.zeros_like)

    def from_original_grad(self, x):
        return self._tr(tf.from_original_grad, x, np.ones_like)

    def _get_diff(self, x=None, x_trans=None):
        """log |dx / dx_trans| (density.py:162-171)"""
        if x is not None:
            return -np.sum(np.log(np.abs(self.from_original_grad(x))), axis=-1)
        if x_trans is not None:
            return np.sum(np.log(np.abs(self.to_original_grad(x_trans))), axis=-1)
        raise ValueError('x and x_trans cannot both be None.')

    def to_original_density(self, density, x_trans=None, x=None):
        return np.asarray(density) - self._get_diff(x, x_trans)

    def from_original_density(self, density, x=None, x_trans=None):
        return np.asarray(density) + self._get_diff(x, x_trans)

    # ------------------------------------------------------------------ device
    def __getstate__(self):
        st = dict(self.__dict__)
        st['_handle'] = None
        st['_synced'] = set()
        st['_dirty'] = True
        return st

    def to_spec(self):
        spec = self._surrogate.to_spec()
        ud = bool(self._use_decay and hasattr(self, '_mu') and self._alpha_2 is not None)
        spec['use_decay'] = ud
        if ud:
            spec.update(d_mu=self._mu.copy(), d_hess=self._hess.copy(), d_alpha2=float(self._alpha_2),
                        d_gamma=float(self._gamma))
        if self._input_scales is None:
            spec['transform_ranges'] = None
        else:
            spec['transform_ranges'] = self._input_scales.copy()
            spec['hard_bounds'] = self._hard_bounds.copy()
        if self.likelihood is not None:
            spec['epilogue'] = self.likelihood.to_spec()      # un-whitened (what the oracle restates); _sync whitens
        if getattr(self, '_prior', None) is not None:
            spec['prior'] = self._prior.to_spec()
        return spec

    def _sync(self, original_space=False):
        """device model of the transformed-space density (the sampler's view) or, with original_space=True,
        of the same density without the variable transform (density.py:503-507 `j = np.eye`)."""
        from . import _cabi
        from .runtime import default_device
        sv = self._surrogate.__dict__.get('_version', 0)
        if self._dirty or sv != getattr(self, '_sur_version', None):
            if not self._surrogate._has_coef():
                raise RuntimeError('the surrogate has no coefficients yet: call fit() first.')
            self._synced = set()
            self._dirty = False
            self._sur_version = sv
        if self._handle is None:
            self._handle = {}
            self._synced = set()
        key = bool(original_space)
        if key not in self._handle:
            dev = self._surrogate._device
            self._handle[key] = _cabi.Handle(default_device() if dev is None else dev)
        if key not in self._synced:
            spec = self.to_spec()
            if original_space:
                spec['transform_ranges'] = None
            if self.likelihood is not None:
                spec = whiten_spec(spec, self.likelihood)
            self._handle[key].set_model(spec)
            self._synced.add(key)
        return self._handle[key]

    # ------------------------------------------------------------------ evaluation (density.py:675-754)
    def logp_and_grad(self, x, original_space=None, use_surrogate=None):
        """x: (n,) or (C, n); one kernel launch for all points."""
        if use_surrogate is False:
            raise NotImplementedError('this Density only holds the surrogate.')
        original_space = self.original_space if original_space is None else bool(original_space)
        x = np.asarray(x, dtype=np.float64)
        single = x.ndim == 1
        lp, g = self._sync(original_space).logp_and_grad_batch(np.atleast_2d(x))
        return (lp[0], g[0]) if single else (lp, g)

    def logp(self, x, original_space=None, use_surrogate=None):
        return self.logp_and_grad(x, original_space, use_surrogate)[0]

    __call__ = logp

    def grad(self, x, original_space=None, use_surrogate=None):
        return self.logp_and_grad(x, original_space, use_surrogate)[1]

    # ------------------------------------------------------------------ fit (density.py:813-838)
    def fit(self, x, y=None, comm=None):
        """
        Fit the surrogate.  Either fit(var_dicts) with objects exposing `_fun[name]` like the reference's
        VariableDict, or fit(x, y) with x (N, n) points of the ORIGINAL space and y (N,) / (N, m) outputs
        (column 0 = logp).
        """
        if y is None:
            su = self._surrogate
            x_arr = np.array([np.concatenate([np.atleast_1d(vd._fun[v]) for v in self.input_vars]) for vd in x])
            y = np.array([np.concatenate([np.atleast_1d(vd._fun[v]) for v in su.output_vars]) for vd in x])
            logp = np.array([np.atleast_1d(vd._fun[self.density_name])[0] for vd in x])
            x = x_arr
        else:
            x = np.asarray(x, dtype=np.float64)
            y = np.asarray(y, dtype=np.float64)
            if y.ndim == 1:
                y = y[:, None]
            # logp of the pipeline, for center_max (poly.py:277-286): output #0, or the likelihood of the outputs
            logp = y[:, 0].copy() if self.likelihood is None else self.likelihood.logp(y)
            if getattr(self, '_prior', None) is not None:
                logp = logp + self._prior.logp(x)
        if self._use_decay:
            self._set_decay(x, comm=comm)
        su = self._surrogate
        xs = x
        if su._input_scales is not None:
            xs = (x - su._input_scales[:, 0]) / su._input_scales_diff
        su.fit(xs, y, logp, comm=comm)
        self._dirty = True

    # ------------------------------------------------------------------ adapter
    @classmethod
    def from_reference(cls, ref, device=None, likelihood=None, prior=None):
        """
        Build from a fitted bayesfast.Density (duck-typed: reads _surrogate_list, _module_list, _input_scales,
        _hard_bounds, decay attributes).  Raises if the density is not surrogate-only / PolyModel-representable.
        With `likelihood` (a GaussianLikelihood restating the module after the surrogate) the surrogate must replace
        every module but that one; with `prior` as well (a GaussianPrior restating a further, last module that adds a prior on
        the inputs to the likelihood: module_2 of examples/des-y1-w-cosmosis.ipynb) every module but those two.
        """
        sl = list(getattr(ref, '_surrogate_list', []))
        if len(sl) != 1:
            raise ValueError('the B200 path needs a density with exactly one surrogate, got {}.'.format(len(sl)))
        rs = sl[0]
        if not (hasattr(rs, '_configs') and hasattr(rs, '_recipe')):
            raise ValueError('the surrogate is not a PolyModel.')
        n_mod = len(getattr(ref, '_module_list', []))
        i_step, n_step = rs._scope
        if prior is not None and likelihood is None:
            raise ValueError('a prior module needs the likelihood module it is added to.')
        if likelihood is not None:
            n_tail = 1 + (prior is not None)
            if not (i_step == 0 and n_step == n_mod - n_tail):
                raise ValueError('with a likelihood{} the surrogate must replace all modules but the last {} '
                                 '(scope={}, {} modules).'.format(' and a prior' if prior is not None else '', n_tail,
                                                                  tuple(rs._scope), n_mod))
        elif not (i_step % max(n_mod, 1) == 0 and n_step == n_mod):
            raise ValueError('the surrogate must replace the whole module list (scope={}, {} modules): arbitrary '
                             'Python modules cannot run on the device.'.format(tuple(rs._scope), n_mod))
        if not getattr(ref, 'use_surrogate', True):
            raise ValueError('density.use_surrogate is False: the true model cannot run on the device.')
        cfgs = []
        for c in rs._configs:
            pc = PolyConfig(c.order, np.asarray(c._input_mask), np.asarray(c._output_mask))
            if c._coef is None:
                raise ValueError('the reference surrogate has not been fitted.')
            for i in range(pc.output_size):
                pc._set(pack_dense(c.order, np.asarray(c._coef)[i], pc.input_size), i)
            cfgs.append(pc)
        su = PolyModel(cfgs, bound_options=dict(use_bound=rs._use_bound, alpha=rs._alpha, alpha_p=rs._alpha_p,
                                                center_max=rs._center_max),
                       input_size=rs._input_size, output_size=rs._output_size,
                       input_scales=None if rs._input_scales is None else np.array(rs._input_scales),
                       device=device)
        if rs._use_bound and hasattr(rs, '_mu'):
            su._mu, su._hess, su._f_mu = np.array(rs._mu), np.array(rs._hess), np.atleast_1d(rs._f_mu).astype(float)
        isc = getattr(ref, '_input_scales', None)
        hb = getattr(ref, '_hard_bounds', False)
        den = cls(su, input_scales=None if isc is None else np.array(isc),
                  hard_bounds=hb if isinstance(hb, bool) else np.array(hb),
                  decay_options=dict(use_decay=ref._use_decay, alpha=ref._alpha, alpha_p=ref._alpha_p,
                                     gamma=ref._gamma),
                  density_name=getattr(ref, 'density_name', '__var__'), likelihood=likelihood, prior=prior)
        if ref._use_decay and hasattr(ref, '_mu'):
            den._mu, den._hess = np.array(ref._mu), np.array(ref._hess)
            den._alpha_2 = float(ref._alpha_2)
        return den
