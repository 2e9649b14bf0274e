"""
Host orchestration of PolyModel.fit / _set_bound on the device (reference: bayesfast/modules/poly.py:505-589,
262-292; bayesfast/core/density.py:796-811).

  rows (x, y, w) --H2D--> fused feature expansion + FP64 DMMA Gram (+ shifted moments)   [csrc/bfb_fit.cu]
                 --(NCCL all-reduce of ONE packed buffer when rows are sharded over GPUs)-->
                 equilibrated Cholesky solve + refinement on every rank (bit-identical inputs -> identical
                 coefficients, no broadcast needed) --> packed coefficients --> PolyConfig._set
"""
import ctypes as C

import numpy as np

from . import _cabi
from .runtime import dist_info, allreduce_values, broadcast_vector, pick_center


def _allreduce_buffer(h, group, stats=None):
    """The exchange step of a sharded fit: the packed partial sums (upper block triangle of the Gram, X^T y, shifted moments,
    row count: bfb_fit_exchange_pack) are all-reduced in place -- NCCL on the device buffer (zero-copy torch view), or staged
    through the host for a backend without device collectives (gloo) -- and unpacked.  stats: dict receiving bytes / seconds."""
    import time
    import torch
    import torch.distributed as dist
    L = _cabi.lib()
    ptr, n = C.c_void_p(), C.c_int64(0)
    _cabi.check(L.bfb_fit_exchange_pack(h._h, C.byref(ptr), C.byref(n)))
    n = int(n.value)
    t0 = time.perf_counter()
    if dist.get_backend(group) == 'nccl':
        t = _wrap_device_ptr(ptr.value, n, h.device)
        dist.all_reduce(t, group=group)
        torch.cuda.synchronize(h.device)
    else:
        host = np.empty(n)
        _cabi.check(L.bfb_fit_exchange_host(h._h, host.ctypes.data, 0))
        t = torch.from_numpy(host)
        dist.all_reduce(t, group=group)
        _cabi.check(L.bfb_fit_exchange_host(h._h, host.ctypes.data, 1))
    dt = time.perf_counter() - t0
    _cabi.check(L.bfb_fit_exchange_unpack(h._h))
    if stats is not None:
        stats.update(allreduce_bytes=8 * n, allreduce_s=dt, backend=dist.get_backend(group))


def _common_shift(x, n, rank, world, group, device):
    """reference point of the shifted moments, identical on all ranks: the first row of rank 0 (zeros if it has none)"""
    shift = np.ascontiguousarray(x[0] if x.shape[0] else np.zeros(n), dtype=np.float64)
    if world > 1:
        shift = np.ascontiguousarray(broadcast_vector(shift if rank == 0 else np.zeros(n), 0, group, device))
    return shift


def _wrap_device_ptr(ptr, n, device):
    """zero-copy torch view of `n` doubles at device address `ptr` (via __cuda_array_interface__)"""
    import torch

    class _Arr:
        __cuda_array_interface__ = {'shape': (n,), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 3,
                                    'strides': None}
    return torch.as_tensor(_Arr(), device='cuda:{}'.format(device))


def _handle_for(model):
    h = model._dev()
    # the configs must be on the device before bfb_fit_begin; coefficients may still be missing (zeros)
    h.set_model(model.to_spec(with_bound=False))
    model._dirty = True
    return h


def _is_cuda(t):
    return hasattr(t, 'is_cuda') and t.is_cuda


def _fit_polymodel_device(model, x, y, logp, w, comm):
    """PolyModel.fit on DEVICE-RESIDENT rows (CUDA torch tensors, float64): the samples of a run that stayed on the GPU feed the
    next surrogate fit without a host round trip (Recipe._sam_step, core/recipe.py:1074-1157); only the n-vector shift, the
    center_max row and the packed coefficients cross PCIe."""
    import torch
    n, m = model._input_size, model._output_size
    if not (x.dim() == 2 and x.shape[1] == n and y.dim() == 2 and y.shape[1] == m and x.shape[0] == y.shape[0]):
        raise ValueError('x should have shape (N, {}) and y (N, {}).'.format(n, m))
    x, y = x.contiguous().double(), y.contiguous().double()
    rank, world, group = dist_info(comm)
    h = _handle_for(model)
    if x.device.index != h.device:
        raise ValueError('the rows live on cuda:{} but the model on cuda:{}.'.format(x.device.index, h.device))
    n_total = x.shape[0]
    if world > 1:
        n_total = int(allreduce_values([n_total], 'sum', group, h.device, 'int64')[0])
    if n_total < model.n_param:
        raise ValueError('I need at least {} points, but you only gave me {}.'.format(model.n_param, n_total))
    wd = None
    if w is not None:
        wd = (w if _is_cuda(w) else torch.as_tensor(np.asarray(w, dtype=np.float64), device=x.device)).contiguous().double()
        if wd.shape != (x.shape[0],):
            raise ValueError('invalid shape for w.')
    L = _cabi.lib()
    x0 = x[0].cpu().numpy() if x.shape[0] else np.zeros(n)
    shift = _common_shift(x0[None] if x.shape[0] else np.zeros((0, n)), n, rank, world, group, h.device)
    torch.cuda.current_stream(x.device).synchronize()
    _cabi.check(L.bfb_fit_begin(h._h, shift.ctypes.data))
    _cabi.check(L.bfb_fit_accumulate(h._h, x.data_ptr(), y.data_ptr(), None if wd is None else wd.data_ptr(), x.shape[0],
                                     _cabi.BFB_DEVICE))
    model._fit_kernel_ms = h.last_kernel_ms()
    if world > 1:
        model._fit_exchange = {}
        _allreduce_buffer(h, group, model._fit_exchange)
    total = sum(_cabi.n_packed(c.order, c.input_size) * c.output_size for c in model._configs)
    coef = np.empty(total)
    rr = C.c_double(0.)
    _cabi.check(L.bfb_fit_solve(h._h, coef.ctypes.data_as(_cabi._dp), C.byref(rr)))
    model._fit_rel_resid = float(rr.value)
    off, packed = 0, []
    for c in model._configs:
        k = _cabi.n_packed(c.order, c.input_size)
        packed.append(coef[off:off + k * c.output_size].reshape(c.output_size, k))
        off += k * c.output_size
    model._install(packed)
    if model._use_bound and not model._all_linear:
        set_bound(model, x, logp, _handle=h, _have_moments=True, comm=comm)


def fit_polymodel(model, x, y, logp=None, w=None, comm=None, refine=1):
    if _is_cuda(x):
        return _fit_polymodel_device(model, x, y, logp, w, comm)
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n, m = model._input_size, model._output_size
    if not (x.ndim == 2 and x.shape[-1] == n):
        raise ValueError('x should be a 2-d array, with shape (# of points, # of input_size), instead of '
                         '{}.'.format(x.shape))
    if not (y.ndim == 2 and y.shape[-1] == m):
        raise ValueError('y should be a 2-d array, with shape (# of points, # of output_size), instead of '
                         '{}.'.format(y.shape))
    if x.shape[0] != y.shape[0]:
        raise ValueError('x and y have different # of points.')
    rank, world, group = dist_info(comm)
    n_total = x.shape[0]
    if world > 1:
        n_total = int(allreduce_values([n_total], 'sum', group, model._dev().device, 'int64')[0])
    if n_total < model.n_param:
        raise ValueError('I need at least {} points, but you only gave me {}.'.format(model.n_param, n_total))
    if w is not None:
        w = np.atleast_1d(np.asarray(w, dtype=np.float64))
        if not (w.ndim == 1 and w.shape[0] == x.shape[0]):
            raise ValueError('invalid shape for w.')
    h = _handle_for(model)
    L = _cabi.lib()
    shift = _common_shift(x, n, rank, world, group, h.device)
    _cabi.check(L.bfb_fit_begin(h._h, shift.ctypes.data))
    xc, yc = _cabi.f64(x), _cabi.f64(y)
    wc = None if w is None else _cabi.f64(w)
    _cabi.check(L.bfb_fit_accumulate(h._h, xc.ctypes.data, yc.ctypes.data, None if wc is None else wc.ctypes.data,
                                     xc.shape[0], _cabi.BFB_HOST))
    model._fit_kernel_ms = h.last_kernel_ms()
    if world > 1:
        model._fit_exchange = {}
        _allreduce_buffer(h, group, model._fit_exchange)
    total = sum(_cabi.n_packed(c.order, c.input_size) * c.output_size for c in model._configs)
    coef = np.empty(total)
    rr = C.c_double(0.)
    _cabi.check(L.bfb_fit_solve(h._h, coef.ctypes.data_as(_cabi._dp), C.byref(rr)))
    model._fit_rel_resid = float(rr.value)
    off, packed = 0, []
    for c in model._configs:
        k = _cabi.n_packed(c.order, c.input_size)
        packed.append(coef[off:off + k * c.output_size].reshape(c.output_size, k))
        off += k * c.output_size
    model._install(packed)
    if model._use_bound and not model._all_linear:
        set_bound(model, x, logp, _handle=h, _have_moments=True, comm=comm)


def ellipsoid(model, x, alpha_p, _handle=None, _have_moments=False, comm=None):
    """mean, inverse covariance and radius alpha of the points x (poly.py:266-276, density.py:802-811)"""
    rank, world, group = dist_info(comm)
    L = _cabi.lib()
    h = _handle
    n = x.shape[1]
    if h is None:
        h = _handle_for(model)
    if not _have_moments:
        if x.shape[1] != model._input_size:
            raise ValueError('invalid value for x.')
        shift = _common_shift(x, n, rank, world, group, h.device)
        _cabi.check(L.bfb_fit_begin(h._h, shift.ctypes.data))
        xc = _cabi.f64(x)
        yz = np.zeros((xc.shape[0], model._output_size))
        _cabi.check(L.bfb_fit_accumulate(h._h, xc.ctypes.data, yz.ctypes.data, None, xc.shape[0], _cabi.BFB_HOST))
        if world > 1:
            _allreduce_buffer(h, group)
    mu, cov = np.empty(n), np.empty((n, n))
    _cabi.check(L.bfb_fit_moments(h._h, mu.ctypes.data_as(_cabi._dp), cov.ctypes.data_as(_cabi._dp)))
    hess = np.linalg.inv(cov)                        # n x n host glue, the same LAPACK call as poly.py:268
    alpha = None
    if alpha_p is not None:
        mb = C.c_double(0.)
        want_all = alpha_p < 100.
        if _is_cuda(x):
            import torch
            xc = x.contiguous().double()
            bt = torch.empty(xc.shape[0], dtype=torch.float64, device=xc.device) if want_all else None
            torch.cuda.current_stream(xc.device).synchronize()
            _cabi.check(L.bfb_fit_max_beta(h._h, xc.data_ptr(), xc.shape[0], mu.ctypes.data_as(_cabi._dp),
                                           hess.ctypes.data_as(_cabi._dp), C.byref(mb), None if bt is None else bt.data_ptr(),
                                           _cabi.BFB_DEVICE))
            beta = None if bt is None else bt.cpu().numpy()
        else:
            xc = _cabi.f64(x)
            beta = np.empty(xc.shape[0]) if want_all else None
            _cabi.check(L.bfb_fit_max_beta(h._h, xc.ctypes.data, xc.shape[0], mu.ctypes.data_as(_cabi._dp),
                                           hess.ctypes.data_as(_cabi._dp), C.byref(mb),
                                           None if beta is None else beta.ctypes.data, _cabi.BFB_HOST))
        if want_all:
            if world > 1:
                raise NotImplementedError('alpha_p < 100 (a percentile of the radii) is not supported with sharded '
                                          'rows; use alpha_p >= 100.')
            alpha = float(np.percentile(beta, alpha_p))       # poly.py:273
        else:
            mx = float(mb.value) if xc.shape[0] else 0.
            if world > 1:
                mx = float(allreduce_values([mx], 'max', group, h.device)[0])
            alpha = mx * alpha_p / 100.                       # poly.py:275
    return mu, hess, alpha


def set_bound(model, x, logp=None, _handle=None, _have_moments=False, comm=None):
    """PolyModel._set_bound (poly.py:262-292)"""
    from .poly import warn_center_max
    if _is_cuda(x):
        if not (x.dim() == 2 and x.shape[-1] == model._input_size):
            raise ValueError('invalid value for x.')
    else:
        x = np.ascontiguousarray(x, dtype=np.float64)
        if not (x.ndim == 2 and x.shape[-1] == model._input_size):
            raise ValueError('invalid value for x.')
    rank, world, group = dist_info(comm)
    mu, hess, alpha = ellipsoid(model, x, model._alpha_p, _handle, _have_moments, comm)
    model._mu, model._hess = mu, hess
    if model._alpha_p is not None:
        model._alpha = alpha
    mu_f = mu
    if model._center_max:
        # the local candidate is validated BEFORE any collective, and every rank always takes part in it (a rank without
        # rows or with an invalid logp contributes -inf), so the ranks can neither hang nor pick different centres
        best, x_best = -np.inf, np.zeros(model._input_size)
        try:
            if _is_cuda(logp):                              # device-resident logp: only the winning row comes to the host
                assert logp.dim() == 1 and logp.shape[0] == x.shape[0]
                lp = logp
                i = int(lp.argmax()) if x.shape[0] > 0 else 0
            else:
                lp = np.asarray(logp, dtype=np.float64)
                assert lp.ndim == 1 and lp.shape[0] == x.shape[0]
                i = int(np.argmax(lp)) if x.shape[0] > 0 else 0                      # poly.py:281
            if x.shape[0] > 0:
                best = float(lp[i])
                x_best = x[i].cpu().numpy() if _is_cuda(x) else x[i]
                if world > 1 and not np.isfinite(best):     # a NaN / inf maximum cannot win the vote over the ranks
                    best = -np.inf
            valid = x.shape[0] > 0 or world > 1
        except Exception:
            valid = False
        if world > 1:
            picked = pick_center(best if valid else -np.inf, x_best, group, model._dev().device)
        else:
            picked = x_best if valid else None
        if picked is None:
            warn_center_max()
        else:
            mu_f = picked
    # f_mu is evaluated with the bound disabled (poly.py:288-292)
    save = model._use_bound
    try:
        model._use_bound = False
        model._dirty = True
        model._f_mu = model._fun(np.ascontiguousarray(mu_f))
    finally:
        model._use_bound = save
        model._dirty = True
