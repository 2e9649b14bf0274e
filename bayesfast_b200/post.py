"""
The steps on either side of the sampler inside the reference's Recipe (core/recipe.py:1074-1157 `_sam_step`, :1286-1297
`_pos_step`) on the device: importance weights with truncation, and SystematicResampler (utils/misc.py:21-110) whose argsort of
the previous density values runs as a bitonic network on the GPU (csrc/bfb_post.cu).  Inputs may be numpy arrays (copied in and
out) or CUDA torch tensors (device-resident: nothing crosses PCIe but the n resampled indices / the five summary numbers).
"""
import ctypes as C
import warnings

import numpy as np

from . import _cabi

__all__ = ['SystematicResampler', 'importance_weights']


def _is_cuda(t):
    return hasattr(t, 'is_cuda') and t.is_cuda


def _handle(device=None, handle=None):
    if handle is not None:
        return handle
    from .runtime import default_device
    return _cabi.Handle(default_device() if device is None else device)


def _lib():
    L = _cabi.lib()
    L.bfb_importance_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int]
    L.bfb_argsort_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    return L


def importance_weights(logp, logq, k_trunc=0.25, handle=None):
    """
    PostStep's importance weights (core/recipe.py:1286-1297): weights = exp(logp - logq), weights_trunc = clip(weights, 0,
    mean(weights) * n ** k_trunc) (k_trunc < 0: no truncation).  Returns (weights, weights_trunc, stats) with stats = dict(sum,
    cap, sum_trunc, sum_trunc_sq, max_trunc, n_eff); the arrays are of the kind passed in (numpy or CUDA torch).
    """
    L = _lib()
    if _is_cuda(logp):
        import torch
        lp, lq = logp.contiguous().double(), logq.contiguous().double()
        h = _handle(lp.device.index, handle)
        w, wt = torch.empty_like(lp), torch.empty_like(lp)
        st = np.empty(5)
        torch.cuda.current_stream(lp.device).synchronize()
        _cabi.check(L.bfb_importance_weights(h._h, lp.data_ptr(), lq.data_ptr(), lp.numel(), float(k_trunc), w.data_ptr(),
                                             wt.data_ptr(), st.ctypes.data, _cabi.BFB_DEVICE))
    else:
        lp, lq = _cabi.f64(logp).reshape(-1), _cabi.f64(logq).reshape(-1)
        if lp.shape != lq.shape:
            raise ValueError('logp and logq should have the same shape.')
        h = _handle(None, handle)
        w, wt, st = np.empty_like(lp), np.empty_like(lp), np.empty(5)
        _cabi.check(L.bfb_importance_weights(h._h, lp.ctypes.data, lq.ctypes.data, lp.size, float(k_trunc), w.ctypes.data,
                                             wt.ctypes.data, st.ctypes.data, _cabi.BFB_HOST))
    stats = dict(sum=st[0], cap=st[1], sum_trunc=st[2], sum_trunc_sq=st[3], max_trunc=st[4],
                 n_eff=st[2]**2 / st[3] if st[3] > 0 else 0.)
    return w, wt, stats


class SystematicResampler:
    """
    Systematically resamples the input array (utils/misc.py:21-110, same parameters and errors): the returned indices are
    argsort(a)[i_all] with i_all the systematic positions between the percentile nodes.
    """

    def __init__(self, nodes=(1., 100.), weights=None, require_unique=True, handle=None):
        try:
            self._nodes = np.asarray(nodes, dtype=np.float64)
            assert self._nodes.ndim == 1 and self._nodes.size > 1
            assert np.all(np.diff(self._nodes) > 0)
            assert self._nodes[0] >= 0 and self._nodes[-1] <= 100
            self._n_node = self._nodes.size
        except Exception:
            raise ValueError('invalid value for nodes.')
        if weights is None:
            self._weights = np.ones(self._n_node - 1) / (self._n_node - 1)
        else:
            try:
                self._weights = np.asarray(weights, dtype=np.float64)
                assert np.all(self._weights) > 0
                assert self._weights.ndim == 1
                assert self._weights.size == self._n_node - 1
                self._weights = self._weights / np.sum(self._weights)
            except Exception:
                raise ValueError('invalid value for weights.')
        self._require_unique = bool(require_unique)
        self._handle = handle

    def positions(self, m, n):
        """i_all of utils/misc.py:88-101 for an array of length m: the host part (n numbers)"""
        n_w = (n * self._weights).astype(np.int64)
        n_w[-1] += n - np.sum(n_w)
        n_c = np.cumsum(np.insert(n_w, 0, 0))
        i_all = np.empty(n, dtype=np.int64)
        for j in range(self._n_node - 1):
            ep = (j == self._n_node - 2)
            i_j = np.linspace(self._nodes[j] * (m - 1) / 100, self._nodes[j + 1] * (m - 1) / 100, n_w[j], ep)
            i_all[n_c[j]:n_c[j + 1]] = i_j.astype(np.int64)
        return i_all

    def run(self, a, n):
        L = _lib()
        try:
            n = int(n)
            assert n > 0
        except Exception:
            raise ValueError('invalid value for n.')
        cuda = _is_cuda(a)
        if cuda:
            if a.dim() != 1:
                raise ValueError('invalid value for a.')
            m = a.numel()
        else:
            try:
                a = np.asarray(a, dtype=np.float64)
                assert a.ndim == 1
            except Exception:
                raise ValueError('invalid value for a.')
            m = a.size
        i_all = self.positions(m, n)
        if np.unique(i_all).size < i_all.size:
            message = ('{:.1f}% of the resampled points are not unique. Please consider giving me more '
                       'points.'.format(100 - np.unique(i_all).size / i_all.size * 100))
            if self._require_unique:
                raise RuntimeError(message)
            warnings.warn(message, RuntimeWarning)
        out = np.empty(n, dtype=np.int64)
        if cuda:
            import torch
            ad = a.contiguous().double()
            h = _handle(ad.device.index, self._handle)
            pos = torch.from_numpy(i_all).to(ad.device)
            o = torch.empty(n, dtype=torch.int64, device=ad.device)
            torch.cuda.current_stream(ad.device).synchronize()
            _cabi.check(L.bfb_argsort_gather(h._h, ad.data_ptr(), m, pos.data_ptr(), n, o.data_ptr(), _cabi.BFB_DEVICE))
            return o
        h = _handle(None, self._handle)
        ac = _cabi.f64(a)
        _cabi.check(L.bfb_argsort_gather(h._h, ac.ctypes.data, m, i_all.ctypes.data, n, out.ctypes.data, _cabi.BFB_HOST))
        return out

    __call__ = run
