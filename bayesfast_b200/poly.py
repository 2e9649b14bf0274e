"""
PolyConfig / PolyModel with the interface of bayesfast.modules.poly (reference: bayesfast/modules/poly.py:19-597,
bayesfast/core/module.py:20-227, 558-687), evaluated and fitted by libbfb200 (sm_100a CUDA) through ctypes.

Differences from the reference that a user can observe:
  * fit() solves equilibrated normal equations (FP64 Gram on the GPU + Cholesky + refinement against the data)
    instead of LAPACK gelsd; coefficients agree to ~1e-10 relative for well-posed fits (tests/test_gpu_fit.py).
  * dense coefficient tensors (`configs[i]._coef`) hold zeros where the reference leaves np.empty garbage.
  * eval_batch(X) evaluates many points in one kernel launch; the single-point methods are thin wrappers.
There is no CPU evaluation path: every numeric method needs the CUDA library.
"""
import warnings
from collections import namedtuple

import numpy as np

from . import _cabi

__all__ = ['PolyConfig', 'PolyModel', 'BoundOptions']

BoundOptions = namedtuple('BoundOptions', ('use_bound', 'alpha', 'alpha_p', 'center_max'))
SurrogateScope = namedtuple('SurrogateScope', ['i_step', 'n_step'])

_ORDERS = ('linear', 'quadratic', 'cubic-2', 'cubic-3')


def _mask(v):
    if v is None:
        return None
    a = np.unique(np.asarray(v, dtype=np.int64))      # sorted + unique, poly.py:56-76
    a.flags.writeable = False
    return a


class PolyConfig:
    """One polynomial block: order + input/output masks + coefficient tensor (poly.py:19-158)."""

    def __init__(self, order, input_mask=None, output_mask=None):
        if order not in _ORDERS:
            raise ValueError('order should be one of ("linear", "quadratic", "cubic-2", "cubic-3"), '
                             'instead of "{}".'.format(order))
        self._order = order
        self._input_mask = _mask(input_mask)
        self._output_mask = _mask(output_mask)
        self._coef = None
        self._packed = None     # (output_size, n_packed): the independent coefficients, lstsq column order

    order = property(lambda self: self._order)
    input_mask = property(lambda self: self._input_mask)
    output_mask = property(lambda self: self._output_mask)

    def _set_input_mask(self, im):
        self._input_mask = _mask(im)

    def _set_output_mask(self, om):
        self._output_mask = _mask(om)

    @property
    def input_size(self):
        return None if self._input_mask is None else self._input_mask.size

    @property
    def output_size(self):
        return None if self._output_mask is None else self._output_mask.size

    def _need_masks(self):
        if self._input_mask is None or self._output_mask is None:
            raise RuntimeError('you have not defined self.input_mask and/or self.output_mask yet.')

    @property
    def _A_shape(self):
        """shape of the dense evaluation tensor (poly.py:87-108)"""
        self._need_masks()
        no, ni = self.output_size, self.input_size
        return {'linear': (no, ni + 1), 'quadratic': (no, ni, ni), 'cubic-2': (no, ni, ni),
                'cubic-3': (no, ni, ni, ni)}[self._order]

    @property
    def _a_shape(self):
        """number of independent coefficients per output (poly.py:110-129)"""
        self._need_masks()
        return (_cabi.n_packed(self._order, self.input_size),)

    def _set(self, a, i):
        """install the independent coefficients `a` of output #i (poly.py:131-158, _poly.pyx:183-214)"""
        self._need_masks()
        a = np.ascontiguousarray(a, dtype=np.float64)
        i = int(i)
        if a.shape != self._a_shape:
            raise ValueError('shape of a {} does not match the expected shape {}.'.format(a.shape, self._a_shape))
        if not 0 <= i < self.output_size:
            raise ValueError('i = {} out of range for self.output_size = {}.'.format(i, self.output_size))
        if self._packed is None:
            self._packed = np.zeros((self.output_size,) + self._a_shape)
        if self._coef is None:
            self._coef = np.zeros(self._A_shape)
        self._packed[i] = a
        self._coef[i] = unpack_dense(self._order, a, self.input_size)


def unpack_dense(order, a, n):
    """packed independent coefficients -> dense evaluation tensor of the reference (zeros elsewhere)."""
    if order == 'linear':
        return np.array(a, dtype=np.float64)
    if order == 'quadratic':
        c = np.zeros((n, n))
        c[np.triu_indices(n)] = a
        return c
    if order == 'cubic-2':
        return np.array(a, dtype=np.float64).reshape(n, n)
    c = np.zeros((n, n, n))
    j, k, l = _c3_indices(n)
    c[j, k, l] = a
    return c


def pack_dense(order, c, n):
    """inverse of unpack_dense (reads only the entries the reference's kernels read)."""
    c = np.asarray(c, dtype=np.float64)
    if order == 'linear':
        return c.copy()
    if order == 'quadratic':
        return c[np.triu_indices(n)].copy()
    if order == 'cubic-2':
        return c.reshape(n * n).copy()
    j, k, l = _c3_indices(n)
    return c[j, k, l].copy()


def _c3_indices(n):
    idx = np.array([(j, k, l) for j in range(n) for k in range(j + 1, n) for l in range(k + 1, n)],
                   dtype=np.int64).reshape(-1, 3)
    return idx[:, 0], idx[:, 1], idx[:, 2]


class PolyModel:
    """
    Polynomial surrogate up to cubic order, same constructor and methods as the reference's PolyModel
    (poly.py:161-597) plus `eval_batch`.  `device` selects the GPU (default: LOCAL_RANK or 0).
    """

    def __init__(self, configs, bound_options=None, input_size=None, output_size=None, input_vars='__var__',
                 output_vars='__var__', input_scales=None, scope=(0, 1), fit_options=None, label=None,
                 device=None):
        try:
            self._input_size, self._output_size = int(input_size), int(output_size)
            assert self._input_size > 0 and self._output_size > 0
        except Exception:
            raise ValueError('input_size and output_size should be positive ints.')
        self.input_vars = [input_vars] if isinstance(input_vars, str) else list(input_vars)
        self.output_vars = [output_vars] if isinstance(output_vars, str) else list(output_vars)
        self.label = label
        self.input_scales = input_scales
        try:
            i_step, n_step = scope
            assert n_step > 0
            self._scope = SurrogateScope(int(i_step), int(n_step))
        except Exception:
            raise ValueError('invalid value for scope.')
        self.fit_options = {} if fit_options is None else dict(fit_options)
        if isinstance(configs, str):
            if configs not in _ORDERS:
                raise ValueError('if configs is a str, it should be "linear", "quadratic", "cubic-2" or "cubic-3".')
            configs = list(_ORDERS[:_ORDERS.index(configs) + 1])      # poly.py:182-193
        if isinstance(configs, PolyConfig):
            configs = [configs]
        if not hasattr(configs, '__iter__'):
            raise ValueError('invalid value for configs.')
        cs = []
        for i, conf in enumerate(configs):
            if isinstance(conf, str):
                conf = PolyConfig(conf)
            if not isinstance(conf, PolyConfig):
                raise ValueError('invalid value for the #{} element of configs.'.format(i))
            if conf._input_mask is None:
                conf._set_input_mask(np.arange(self._input_size))
            if conf._output_mask is None:
                conf._set_output_mask(np.arange(self._output_size))
            if conf._input_mask[-1] >= self._input_size or conf._output_mask[-1] >= self._output_size:
                raise ValueError('mask of PolyConfig #{} out of range.'.format(i))
            cs.append(conf)
        self._configs = tuple(cs)
        self._build_recipe()
        if bound_options is None:
            bound_options = {}
        if not isinstance(bound_options, dict):
            raise ValueError('bound_options should be a dict.')
        self.set_bound_options(**bound_options)
        self._device = device
        self._handle = None
        self._dirty = True
        self.reset_counter()

    # ------------------------------------------------------------------ bookkeeping
    @property
    def _dirty(self):
        """True when this object's own device handle does not hold the current model."""
        return self.__dict__.get('_dirty_flag', True)

    @_dirty.setter
    def _dirty(self, v):
        self.__dict__['_dirty_flag'] = bool(v)
        if v:   # every modification bumps the version that dependants (Density) watch
            self.__dict__['_version'] = self.__dict__.get('_version', 0) + 1

    configs = property(lambda self: self._configs)
    n_config = property(lambda self: len(self._configs))
    input_size = property(lambda self: self._input_size)
    output_size = property(lambda self: self._output_size)
    scope = property(lambda self: self._scope)
    recipe = property(lambda self: self._recipe)

    @property
    def input_scales(self):
        return self._input_scales

    @input_scales.setter
    def input_scales(self, scales):
        if scales is None:
            self._input_scales = None
            self._input_scales_diff = 1.
            return
        try:
            scales = np.ascontiguousarray(scales, dtype=np.float64)
            if scales.ndim == 1:
                scales = np.array((np.zeros_like(scales), scales)).T.copy()
            assert scales.ndim == 2 and scales.shape == (self._input_size, 2)
        except Exception:
            raise ValueError('invalid value for input_scales.')
        self._input_scales = scales
        self._input_scales_diff = scales[:, 1] - scales[:, 0]
        self._dirty = True

    @property
    def bound_options(self):
        return BoundOptions(self._use_bound, self._alpha, self._alpha_p, self._center_max)

    def set_bound_options(self, use_bound=True, alpha=None, alpha_p=100., center_max=True):
        """linear extrapolation outside the training ellipsoid (poly.py:234-260)"""
        self._use_bound = bool(use_bound)
        if alpha is None:
            self._alpha = None
        else:
            try:
                alpha = float(alpha)
                assert alpha > 0
            except Exception:
                raise ValueError('invalid value for alpha.')
            self._alpha = alpha
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
        self._center_max = bool(center_max)
        self._dirty = True

    def _build_recipe(self):
        """which config serves which (output, order) pair; one config per pair at most (poly.py:298-337)"""
        rr = np.full((self._output_size, 4), -1)
        names = ('linear', 'quadratic', 'cubic_2', 'cubic_3')
        for ii, conf in enumerate(self._configs):
            col = _ORDERS.index(conf.order)
            if np.any(rr[conf._output_mask, col] >= 0):
                raise ValueError('multiple {} PolyConfig(s) share at least one common output variable. Please '
                                 'check your PolyConfig #{}.'.format(names[col], ii))
            rr[conf._output_mask, col] = ii
        if np.any(np.all(rr < 0, axis=1)):
            raise ValueError('no PolyConfig has output for variable(s) {}.'.format(
                np.flatnonzero(np.all(rr < 0, axis=1))))
        rr.flags.writeable = False
        self._recipe = rr

    @property
    def n_param(self):
        return int(np.sum([conf._a_shape[0] for conf in self._configs]))

    @property
    def _all_linear(self):
        return all(conf.order == 'linear' for conf in self._configs)

    def reset_counter(self):
        self._ncall_fun = self._ncall_jac = self._ncall_fun_and_jac = 0

    ncall_fun = property(lambda self: self._ncall_fun)
    ncall_jac = property(lambda self: self._ncall_jac)
    ncall_fun_and_jac = property(lambda self: self._ncall_fun_and_jac)

    # ------------------------------------------------------------------ device plumbing
    def __getstate__(self):
        st = dict(self.__dict__)
        st['_handle'] = None            # device handles are per process: objects stay picklable / deep-copyable
        st['_dirty_flag'] = True
        return st

    def _dev(self):
        if self._handle is None:
            from .runtime import default_device
            self._handle = _cabi.Handle(default_device() if self._device is None else self._device)
            self._dirty = True
        return self._handle

    def _has_coef(self):
        return all(c._packed is not None for c in self._configs)

    def _bound_active(self):
        return bool(self._use_bound and not self._all_linear and hasattr(self, '_mu') and self._alpha is not None)

    def to_spec(self, with_scales=True, with_bound=True):
        """model interchange dict (see oracle/bf_oracle.py for the field list); 'packed' feeds the C ABI,
        'coef' (dense reference layout) feeds the oracle."""
        spec = dict(n=self._input_size, m=self._output_size, configs=[])
        for c in self._configs:
            spec['configs'].append(dict(order=c.order, input_mask=np.asarray(c._input_mask),
                                        output_mask=np.asarray(c._output_mask),
                                        packed=None if c._packed is None else c._packed.copy(),
                                        coef=None if c._coef is None else c._coef.copy()))
        ub = with_bound and self._bound_active()
        spec['use_bound'] = ub
        if ub:
            spec.update(mu=self._mu.copy(), hess=self._hess.copy(), alpha=float(self._alpha),
                        f_mu=np.atleast_1d(self._f_mu).astype(float))
        spec['input_scales'] = self._input_scales.copy() if (with_scales and self._input_scales is not None) else None
        return spec

    def _sync(self):
        h = self._dev()
        if self._dirty:
            if not self._has_coef():
                raise RuntimeError('the PolyModel has no coefficients yet: call fit() or PolyConfig._set() first.')
            h.set_model(self.to_spec())
            self._dirty = False
        return h

    def mark_dirty(self):
        """call after editing configs[i]._packed / bound attributes by hand"""
        self._dirty = True

    # ------------------------------------------------------------------ evaluation
    def eval_batch(self, X, want_jac=True, rescale=True):
        """
        fun_and_jac for many points in one launch.  X (C, n) -> F (C, m), J (C, m, n).
        rescale=True applies the module-level input_scales like the public fun/jac (module.py:221-227);
        rescale=False is the raw `_fun_and_jac` (poly.py:466).
        """
        X = np.asarray(X, dtype=np.float64)
        if X.ndim != 2 or X.shape[1] != self._input_size:
            raise ValueError('X should have shape (# of points, {}), instead of {}.'.format(
                self._input_size, X.shape))
        if not rescale and self._input_scales is not None:
            # evaluate in the rescaled coordinates: undo the wrapper analytically
            X = X * self._input_scales_diff + self._input_scales[:, 0]
            F, J = self._sync().poly_eval_batch(X, want_jac)
            if J is not None:
                J = J * self._input_scales_diff
            return F, J
        return self._sync().poly_eval_batch(X, want_jac)

    def _x1(self, x):
        x = np.asarray(x, dtype=np.float64)
        if x.shape != (self._input_size,):
            raise ValueError('x should have shape ({},), instead of {}.'.format(self._input_size, x.shape))
        return x[None]

    def _fun(self, x):
        return self.eval_batch(self._x1(x), False, rescale=False)[0][0]

    def _jac(self, x):
        return self.eval_batch(self._x1(x), True, rescale=False)[1][0]

    def _fun_and_jac(self, x):
        F, J = self.eval_batch(self._x1(x), True, rescale=False)
        return F[0], J[0]

    @staticmethod
    def _cat(args):
        return np.concatenate([np.atleast_1d(np.asarray(a, dtype=np.float64)) for a in args])

    def fun(self, *args):
        """public wrapper: concatenates the inputs, returns a list like ModuleBase.fun (module.py:121-147)"""
        self._ncall_fun += 1
        return [self.eval_batch(self._cat(args)[None], False)[0][0]]

    __call__ = fun

    def jac(self, *args):
        self._ncall_jac += 1
        return [self.eval_batch(self._cat(args)[None], True)[1][0]]

    def fun_and_jac(self, *args):
        self._ncall_fun_and_jac += 1
        F, J = self.eval_batch(self._cat(args)[None], True)
        return [F[0]], [J[0]]

    # ------------------------------------------------------------------ fit
    def fit(self, x, y, logp=None, w=None, comm=None, refine=1):
        """
        Least-squares fit of all configs (poly.py:505-589).  x (N, n), y (N, m), optional logp (N,) for
        center_max and row weights w (N,).  With `comm` (a torch.distributed process group or True for the
        default group) x/y/w are this rank's rows and the Gram / moment partial sums are all-reduced over NCCL.
        x / y (/ logp / w) may be CUDA torch tensors: device-resident rows (e.g. samples that never left the GPU) are
        accumulated where they are, only the coefficients come back.
        """
        from .fit import fit_polymodel
        fit_polymodel(self, x, y, logp, w, comm=comm, refine=refine)

    def _set_bound(self, x, logp=None):
        from .fit import set_bound
        set_bound(self, np.asarray(x, dtype=np.float64), logp)

    def _install(self, packed_per_config):
        """packed_per_config[i]: (n_out_i, n_packed_i)"""
        for conf, pk in zip(self._configs, packed_per_config):
            for i in range(conf.output_size):
                conf._set(pk[i], i)
        self._dirty = True


def warn_center_max():
    warnings.warn('invalid value for logp. Disabling center_max for now.', RuntimeWarning)
