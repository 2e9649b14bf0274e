"""
Bounded <-> unbounded variable transform on the host (numpy), same maps as the reference's
bayesfast/transforms/_constraint.pyx:19-221 and its front-end core/density.py:92-157.  Used for x_0
(from_original) and for reporting samples in the original space; inside the sampler the same maps run
on the device (csrc/bfb_eval.cuh: to_original_1).
"""
import numpy as np

__all__ = ['check_scales', 'check_bounds', 'from_original', 'to_original', 'to_original_grad', 'to_original_grad2',
           'from_original_grad']


def check_scales(scales):
    """density.py:39-50"""
    try:
        scales = np.ascontiguousarray(scales, dtype=np.float64)
        if scales.ndim == 1:
            scales = np.array((np.zeros_like(scales), scales)).T.copy()
        if not (scales.ndim == 2 and scales.shape[-1] == 2):
            raise ValueError
    except Exception:
        raise ValueError('Invalid value for input_scales.')
    return scales


def check_bounds(bounds, n):
    """density.py:64-75; returns (n, 2) uint8"""
    if isinstance(bounds, (bool, np.bool_)):
        return np.full((n, 2), int(bool(bounds)), dtype=np.uint8)
    try:
        bounds = np.atleast_1d(bounds).astype(bool).astype(np.uint8).copy()
        if bounds.ndim == 1:
            bounds = np.array((bounds, bounds)).T.copy()
        if not (bounds.ndim == 2 and bounds.shape == (n, 2)):
            raise ValueError
    except Exception:
        raise ValueError('Invalid value for hard_bounds')
    return bounds


def _kinds(hb):
    lo, hi = hb[:, 0].astype(bool), hb[:, 1].astype(bool)
    return lo & hi, lo & ~hi, ~lo & hi


def from_original(x, ranges, hb):
    """_from_original_f, _constraint.pyx:19-38"""
    x = np.asarray(x, dtype=np.float64)
    t = (x - ranges[:, 0]) / (ranges[:, 1] - ranges[:, 0])
    both, lower, upper = _kinds(hb)
    bad = (both & ((t <= 0.) | (t >= 1.))) | (lower & (t <= 0.)) | (upper & (t >= 1.))
    if np.any(bad):
        raise ValueError('variable #{} out of bound.'.format(int(np.argwhere(bad)[0][-1])))
    out = t.copy()
    with np.errstate(all='ignore'):
        out = np.where(both, np.log(t / (1. - t)), out)
        out = np.where(lower, np.log(t), out)
        out = np.where(upper, np.log(1. - t), out)
    return out


def from_original_grad(x, ranges, hb):
    """_from_original_j, _constraint.pyx:53-77"""
    x = np.asarray(x, dtype=np.float64)
    w = ranges[:, 1] - ranges[:, 0]
    t = (x - ranges[:, 0]) / w
    both, lower, upper = _kinds(hb)
    out = np.ones_like(t)
    with np.errstate(all='ignore'):
        out = np.where(both, 1. / t / (1. - t), out)
        out = np.where(lower, 1. / t, out)
        out = np.where(upper, 1. / (t - 1.), out)
    return out / w


def to_original(x, ranges, hb):
    """_to_original_f, _constraint.pyx:133-150"""
    x = np.asarray(x, dtype=np.float64)
    both, lower, upper = _kinds(hb)
    t = x.copy()
    with np.errstate(over='ignore'):
        t = np.where(both, 1. / (1. + np.exp(-x)), t)
        t = np.where(lower, np.exp(x), t)
        t = np.where(upper, 1. - np.exp(x), t)
    return ranges[:, 0] + t * (ranges[:, 1] - ranges[:, 0])


def to_original_grad(x, ranges, hb):
    """_to_original_j, _constraint.pyx:167-186"""
    x = np.asarray(x, dtype=np.float64)
    both, lower, upper = _kinds(hb)
    t = np.ones_like(x)
    with np.errstate(over='ignore'):
        s = 1. / (1. + np.exp(-x))
        t = np.where(both, s * (1. - s), t)
        t = np.where(lower, np.exp(x), t)
        t = np.where(upper, -np.exp(x), t)
    return t * (ranges[:, 1] - ranges[:, 0])


def to_original_grad2(x, ranges, hb):
    """_to_original_jj, _constraint.pyx:203-221"""
    x = np.asarray(x, dtype=np.float64)
    both, lower, upper = _kinds(hb)
    t = np.zeros_like(x)
    with np.errstate(over='ignore', invalid='ignore'):
        e = np.exp(x)
        t = np.where(both, -e * (e - 1.) / (e + 1.) / (e + 1.) / (e + 1.), t)
        t = np.where(lower, e, t)
        t = np.where(upper, -e, t)
    return t * (ranges[:, 1] - ranges[:, 0])
