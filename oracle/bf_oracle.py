"""
CPU ORACLE for bayesfast_b200 -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/libbf_oracle.so (C restatement of the reference's PolyModel evaluation
and NUTS/HMC sampler, see bf_oracle.h) plus a numpy/scipy restatement of PolyModel.fit /
_set_bound (bayesfast/modules/poly.py:505-589, 262-292), which calls the same third-party routine
as the reference (scipy.linalg.lstsq, LAPACK gelsd; scipy is unpinned in the reference's setup.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Parity pin: tests/golden/*.npz (tests/golden/make_golden.py ran the real reference).

Model interchange format ("spec"): a plain dict of numpy arrays, produced by
bayesfast_b200.PolyModel.to_spec() / Density.to_spec() or written by hand:

    n, m                         input / output size
    configs                      list of dict(order=str, input_mask=int64[], output_mask=int64[], coef=ndarray)
                                 coef in the reference's dense layout (poly.py:87-108)
    use_bound, mu, hess, alpha, f_mu            PolyModel bound (poly.py:262-292); use_bound already
                                                 and-ed with "not all linear"
    input_scales                 None or (n, 2): module-level rescale (core/module.py:47-96)
    use_decay, d_mu, d_hess, d_alpha2, d_gamma  Density decay (core/density.py:740-746, 796-811)
    transform_ranges, hard_bounds               None / (n,2) float, (n,2) uint8: Density.input_scales transform
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ORDER = {'linear': 1, 'quadratic': 2, 'cubic-2': 3, 'cubic-3': 4}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)


class _Config(C.Structure):
    _fields_ = [('order', C.c_int32), ('n_in', C.c_int32), ('n_out', C.c_int32),
                ('in_mask', _lp), ('out_mask', _lp), ('coef', _dp)]


class _PolyModel(C.Structure):
    _fields_ = [('n', C.c_int32), ('m', C.c_int32), ('n_config', C.c_int32), ('configs', C.POINTER(_Config)),
                ('use_bound', C.c_int32), ('mu', _dp), ('hess', _dp), ('alpha', C.c_double), ('f_mu', _dp),
                ('use_scales', C.c_int32), ('s0', _dp), ('sdiff', _dp)]


class _Density(C.Structure):
    _fields_ = [('model', C.POINTER(_PolyModel)), ('use_decay', C.c_int32), ('d_mu', _dp), ('d_hess', _dp),
                ('d_alpha2', C.c_double), ('d_gamma', C.c_double), ('use_transform', C.c_int32),
                ('ranges', _dp), ('hard_bounds', _bp), ('use_epilogue', C.c_int32), ('e_d', _dp), ('e_cinv', _dp),
                ('e_c0', C.c_double), ('use_prior', C.c_int32), ('p_w', _dp), ('p_mu', _dp), ('p_c0', C.c_double)]


class _Cfg(C.Structure):
    _fields_ = [('n_iter', C.c_int32), ('n_warmup', C.c_int32), ('max_treedepth', C.c_int32),
                ('n_int_step', C.c_int32), ('max_change', C.c_double), ('adapt_step_size', C.c_int32),
                ('target_accept', C.c_double), ('gamma', C.c_double), ('k', C.c_double), ('t0', C.c_double),
                ('adapt_metric', C.c_int32), ('initial_weight', C.c_double), ('adapt_window', C.c_int32),
                ('update_window', C.c_int32), ('doubling', C.c_int32), ('seed', C.c_uint64),
                ('n_threads', C.c_int32), ('dense_metric', C.c_int32)]


class _Out(C.Structure):
    _fields_ = [('samples', _dp), ('logp', _dp), ('energy', _dp), ('mean_tree_accept', _dp), ('step_size', _dp),
                ('step_size_bar', _dp), ('energy_change', _dp), ('max_energy_change', _dp),
                ('tree_depth', _ip), ('tree_size', _ip), ('diverging', _ip),
                ('final_step', _dp), ('final_var', _dp), ('n_draws', _lp), ('status', _ip)]


def build(force=False):
    """Compile oracle/libbf_oracle.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, 'libbf_oracle.so')
    src = [os.path.join(_HERE, f) for f in ('bf_oracle.c', 'bf_oracle.h')]
    src.append(os.path.join(_HERE, '..', 'include', 'bfb_rng.h'))
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.check_call(['make', '-C', _HERE, '-B', 'libbf_oracle.so'], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, 'libbf_oracle.so')
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.bfo_poly_eval_batch.argtypes = [C.POINTER(_PolyModel), _dp, C.c_int64, _dp, _dp, C.c_int]
        L.bfo_logp_and_grad_batch.argtypes = [C.POINTER(_Density), _dp, C.c_int64, _dp, _dp, C.c_int]
        for f in (L.bfo_nuts_run, L.bfo_hmc_run):
            f.argtypes = [C.POINTER(_Density), C.POINTER(_Cfg), C.c_int64, C.c_int64, _dp, _dp, _dp, _dp,
                          _dp, _dp, C.c_int64, C.POINTER(_Out)]
            f.restype = C.c_int
        L.bfo_tempered_run.argtypes = [C.POINTER(_Density), C.POINTER(_Density), C.c_double, C.c_int, C.POINTER(_Cfg),
                                       C.c_int64, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int64,
                                       C.POINTER(_Out), _dp, _dp]
        L.bfo_tempered_run.restype = C.c_int
        L.bfo_rng_fill.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64, _dp, _dp]
        L.bfo_philox_raw.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        L.bfo_to_original.argtypes = [_dp, _dp, _bp, C.c_int, _dp, _dp, _dp]
        L.bfo_from_original.argtypes = [_dp, _dp, _bp, C.c_int, _dp]
        L.bfo_from_original.restype = C.c_int
        for f in (L.bfo_lsq_quadratic, L.bfo_lsq_cubic_2, L.bfo_lsq_cubic_3):
            f.argtypes = [_dp, _dp, C.c_int64, C.c_int]
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleDensity:
    """Holds the C structs (and the numpy buffers they point into) for one spec."""

    def __init__(self, spec):
        self.spec = spec
        n, m = int(spec['n']), int(spec['m'])
        self.n, self.m = n, m
        self._keep = []
        cfgs = (_Config * len(spec['configs']))()
        for i, cf in enumerate(spec['configs']):
            im = np.ascontiguousarray(cf['input_mask'], dtype=np.int64)
            om = np.ascontiguousarray(cf['output_mask'], dtype=np.int64)
            co = _f64(cf['coef'])
            self._keep += [im, om, co]
            cfgs[i].order = ORDER[cf['order']]
            cfgs[i].n_in, cfgs[i].n_out = im.size, om.size
            cfgs[i].in_mask = im.ctypes.data_as(_lp)
            cfgs[i].out_mask = om.ctypes.data_as(_lp)
            cfgs[i].coef = _d(co)
        self._cfgs = cfgs
        pm = _PolyModel()
        pm.n, pm.m, pm.n_config = n, m, len(spec['configs'])
        pm.configs = C.cast(cfgs, C.POINTER(_Config))
        pm.use_bound = int(bool(spec.get('use_bound', False)))
        mu = _f64(spec['mu']) if pm.use_bound else np.zeros(n)
        hess = _f64(spec['hess']) if pm.use_bound else np.eye(n)
        f_mu = _f64(np.atleast_1d(spec['f_mu'])) if pm.use_bound else np.zeros(m)
        pm.alpha = float(spec['alpha']) if pm.use_bound else 0.
        sc = spec.get('input_scales', None)
        pm.use_scales = int(sc is not None)
        s0 = _f64(np.asarray(sc)[:, 0]) if sc is not None else np.zeros(n)
        sd = _f64(np.asarray(sc)[:, 1] - np.asarray(sc)[:, 0]) if sc is not None else np.ones(n)
        self._keep += [mu, hess, f_mu, s0, sd]
        pm.mu, pm.hess, pm.f_mu, pm.s0, pm.sdiff = _d(mu), _d(hess), _d(f_mu), _d(s0), _d(sd)
        self._pm = pm
        dn = _Density()
        dn.model = C.pointer(pm)
        dn.use_decay = int(bool(spec.get('use_decay', False)))
        d_mu = _f64(spec['d_mu']) if dn.use_decay else np.zeros(n)
        d_hess = _f64(spec['d_hess']) if dn.use_decay else np.eye(n)
        dn.d_alpha2 = float(spec['d_alpha2']) if dn.use_decay else 0.
        dn.d_gamma = float(spec['d_gamma']) if dn.use_decay else 0.
        tr = spec.get('transform_ranges', None)
        dn.use_transform = int(tr is not None)
        ranges = _f64(tr) if tr is not None else np.zeros((n, 2))
        hb = np.ascontiguousarray(spec['hard_bounds'], dtype=np.uint8) if tr is not None else np.zeros((n, 2), np.uint8)
        self._keep += [d_mu, d_hess, ranges, hb]
        dn.d_mu, dn.d_hess, dn.ranges, dn.hard_bounds = _d(d_mu), _d(d_hess), _d(ranges), hb.ctypes.data_as(_bp)
        ep = spec.get('epilogue', None)       # dict(d [m], cinv [m,m], c0): Gaussian likelihood of the m outputs
        dn.use_epilogue = int(ep is not None)
        e_d = _f64(np.atleast_1d(ep['d'])) if ep is not None else np.zeros(m)
        e_ci = _f64(np.atleast_2d(ep['cinv'])) if ep is not None else np.eye(m)
        assert e_d.shape == (m,) and e_ci.shape == (m, m)
        dn.e_c0 = float(ep.get('c0', 0.)) if ep is not None else 0.
        self._keep += [e_d, e_ci]
        dn.e_d, dn.e_cinv = _d(e_d), _d(e_ci)
        pr = spec.get('prior', None)          # dict(idx, mu, sig, c0): independent Gaussian prior on some original-space inputs
        dn.use_prior = int(pr is not None)
        p_w, p_mu = np.zeros(n), np.zeros(n)
        if pr is not None:
            idx = np.asarray(pr['idx'], dtype=np.int64)
            p_w[idx] = 1. / np.asarray(pr['sig'], dtype=np.float64)**2
            p_mu[idx] = np.asarray(pr['mu'], dtype=np.float64)
            dn.p_c0 = float(pr.get('c0', 0.))
        self._keep += [p_w, p_mu]
        dn.p_w, dn.p_mu = _d(p_w), _d(p_mu)
        self._dn = dn

    # PolyModel.fun_and_jac for a batch of points (module rescale included, no Density wrapper)
    def poly_eval_batch(self, X, n_threads=0):
        X = _f64(X).reshape(-1, self.n)
        F = np.empty((X.shape[0], self.m))
        J = np.empty((X.shape[0], self.m, self.n))
        lib().bfo_poly_eval_batch(C.byref(self._pm), _d(X), X.shape[0], _d(F), _d(J), n_threads)
        return F, J

    # Density.logp_and_grad(x, original_space=False) for a batch
    def logp_and_grad_batch(self, X, n_threads=0):
        X = _f64(X).reshape(-1, self.n)
        lp = np.empty(X.shape[0])
        g = np.empty((X.shape[0], self.n))
        lib().bfo_logp_and_grad_batch(C.byref(self._dn), _d(X), X.shape[0], _d(lp), _d(g), n_threads)
        return lp, g

    def run(self, sampler, cfg, x0, step0, var0, mean0=None, draws_u=None, draws_z=None, chain0=0,
            base=None, logxi=0., u0=None):
        """cfg: dict with the _Cfg fields (missing ones take the reference defaults of
        bayesfast/samplers/sample_trace.py:157-166, 499-512).  sampler 'TNUTS' / 'THMC' (samplers/tnuts.py, thmc.py):
        base = OracleDensity of TNTrace.density_base, logxi, u0 [C]; the result then also has 'u' and 'weight'."""
        n = self.n
        x0 = _f64(x0).reshape(-1, n)
        nc = x0.shape[0]
        c = _Cfg()
        dflt = dict(n_iter=1500, n_warmup=500, max_treedepth=10, n_int_step=32, max_change=1000.,
                    adapt_step_size=1, target_accept=0.8, gamma=0.05, k=0.75, t0=10., adapt_metric=1,
                    initial_weight=10., adapt_window=60, update_window=1, doubling=1, seed=0, n_threads=0,
                    dense_metric=0)
        dflt.update(cfg)
        for k, v in dflt.items():
            setattr(c, k, v)
        step0 = _f64(np.broadcast_to(step0, (nc,)))
        vshape = (nc, n, n) if c.dense_metric else (nc, n)      # dense: covariance matrices (metrics.py:94-132)
        var0 = _f64(np.broadcast_to(var0, vshape))
        mean0 = x0.copy() if mean0 is None else _f64(np.broadcast_to(mean0, (nc, n)))
        ni = c.n_iter
        res = dict(samples=np.zeros((nc, ni, n)), final_step=np.zeros((nc, 4)), final_var=np.zeros(vshape),
                   n_draws=np.zeros(nc, np.int64), status=np.zeros(nc, np.int32))
        for k in ('logp', 'energy', 'mean_tree_accept', 'step_size', 'step_size_bar', 'energy_change',
                  'max_energy_change'):
            res[k] = np.zeros((nc, ni))
        for k in ('tree_depth', 'tree_size', 'diverging'):
            res[k] = np.zeros((nc, ni), np.int32)
        o = _Out()
        for k, _t in _Out._fields_:
            o.__setattr__(k, res[k].ctypes.data_as(_t))
        nrep = 0
        pu = pz = None
        if draws_u is not None:
            draws_u = _f64(draws_u).reshape(nc, -1)
            draws_z = _f64(draws_z).reshape(nc, -1)
            nrep = draws_u.shape[1]
            pu, pz = _d(draws_u), _d(draws_z)
        if sampler.upper() in ('TNUTS', 'THMC'):
            assert base is not None and base.n == n
            u0 = _f64(np.broadcast_to(u0, (nc,)))
            res['u'], res['weight'] = np.zeros((nc, ni)), np.zeros((nc, ni))
            rc = lib().bfo_tempered_run(C.byref(self._dn), C.byref(base._dn), float(logxi), int(sampler.upper() == 'TNUTS'),
                                        C.byref(c), nc, int(chain0), _d(x0), _d(u0), _d(step0), _d(var0), _d(mean0),
                                        pu, pz, nrep, C.byref(o), _d(res['u']), _d(res['weight']))
            assert rc >= 0
            return res
        f = lib().bfo_nuts_run if sampler.upper() == 'NUTS' else lib().bfo_hmc_run
        f(C.byref(self._dn), C.byref(c), nc, int(chain0), _d(x0), _d(step0), _d(var0), _d(mean0), pu, pz, nrep,
          C.byref(o))
        return res


def rng_fill(seed, chain, t0, count):
    u = np.empty(count)
    z = np.empty(count)
    lib().bfo_rng_fill(int(seed), int(chain), int(t0), int(count), _d(u), _d(z))
    return u, z


def philox_raw(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().bfo_philox_raw(*[int(v) for v in ctr], *[int(v) for v in key], out)
    return [int(v) for v in out]


def to_original(x, ranges, hard_bounds):
    x = _f64(x)
    ranges = _f64(ranges)
    hb = np.ascontiguousarray(hard_bounds, dtype=np.uint8)
    f, j, jj = np.empty_like(x), np.empty_like(x), np.empty_like(x)
    lib().bfo_to_original(_d(x), _d(ranges), hb.ctypes.data_as(_bp), x.size, _d(f), _d(j), _d(jj))
    return f, j, jj


def from_original(x, ranges, hard_bounds):
    x = _f64(x)
    ranges = _f64(ranges)
    hb = np.ascontiguousarray(hard_bounds, dtype=np.uint8)
    f = np.empty_like(x)
    bad = lib().bfo_from_original(_d(x), _d(ranges), hb.ctypes.data_as(_bp), x.size, _d(f))
    if bad:
        raise ValueError('variable #{} out of bound.'.format(bad - 1))
    return f


# ------------------------------------------------------------------------------------------------
# PolyModel.fit restated with numpy (poly.py:505-589) -- single shared design matrix per recipe row.
# ------------------------------------------------------------------------------------------------
def n_coef(order, n):
    """poly.py:110-129 (_a_shape)"""
    return {'linear': n + 1, 'quadratic': n * (n + 1) // 2, 'cubic-2': n * n,
            'cubic-3': n * (n - 1) * (n - 2) // 6}[order]


def design_block(order, x):
    """_lsq_* of modules/_poly.pyx:143-177 and the [1, x] block of poly.py:533-540."""
    x = _f64(x)
    rows, n = x.shape
    if order == 'linear':
        return np.concatenate((np.ones((rows, 1)), x), axis=1)
    out = np.empty((rows, n_coef(order, n)))
    f = {'quadratic': lib().bfo_lsq_quadratic, 'cubic-2': lib().bfo_lsq_cubic_2,
         'cubic-3': lib().bfo_lsq_cubic_3}[order]
    f(_d(x), _d(out), rows, n)
    return out


def unpack_coef(order, a, n):
    """_set_* of modules/_poly.pyx:183-214; unused entries are zero (the reference leaves np.empty garbage)."""
    a = np.asarray(a)
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
            nl = n - k - 1
            c[j, k, k + 1:] = a[i:i + nl]
            i += nl
    return c


def fit(configs, n, m, x, y, w=None):
    """
    configs: list of dict(order, input_mask, output_mask).  Returns the list of dense coef tensors
    (one per config, reference layout).  Follows poly.py:529-587 output by output.
    """
    from scipy.linalg import lstsq
    x = _f64(x)
    y = _f64(y)
    coefs = []
    for cf in configs:
        ni, no = len(cf['input_mask']), len(cf['output_mask'])
        shp = {'linear': (no, ni + 1), 'quadratic': (no, ni, ni), 'cubic-2': (no, ni, ni),
               'cubic-3': (no, ni, ni, ni)}[cf['order']]
        coefs.append(np.zeros(shp))
    col = {'linear': 0, 'quadratic': 1, 'cubic-2': 2, 'cubic-3': 3}
    for ii in range(m):
        recipe = [-1, -1, -1, -1]
        for jj, cf in enumerate(configs):
            if ii in list(cf['output_mask']):
                recipe[col[cf['order']]] = jj
        blocks, used = [], []
        for jj in recipe:
            if jj >= 0:
                cf = configs[jj]
                blocks.append(design_block(cf['order'], x[:, np.asarray(cf['input_mask'])]))
                used.append(jj)
        A = np.concatenate(blocks, axis=1)
        b = y[:, ii].copy()
        if w is not None:
            b *= w
            A *= np.asarray(w)[:, None]
        sol = lstsq(A, b)[0]
        p = 0
        for jj, blk in zip(used, blocks):
            cf = configs[jj]
            q = int(np.argwhere(np.asarray(cf['output_mask']) == ii)[0, 0])
            coefs[jj][q] = unpack_coef(cf['order'], sol[p:p + blk.shape[1]], len(cf['input_mask']))
            p += blk.shape[1]
    return coefs


def bound_from_points(x, alpha_p=100.):
    """poly.py:262-276 (_set_bound) / density.py:796-811 (_set_decay): mu, hess, alpha."""
    x = _f64(x)
    mu = np.mean(x, axis=0)
    hess = np.linalg.inv(np.atleast_2d(np.cov(x, rowvar=False)))
    beta = np.einsum('ij,jk,ik->i', x - mu, hess, x - mu) ** 0.5
    if alpha_p < 100.:
        alpha = np.percentile(beta, alpha_p)
    else:
        alpha = np.max(beta) * alpha_p / 100.
    return mu, hess, float(alpha)


# ---------------------------------------------------------------------------------------------
# The steps on either side of the sampler in Recipe, numpy restatements
# ---------------------------------------------------------------------------------------------
def systematic_resample(a, n, nodes=(1., 100.), weights=None):
    """SystematicResampler.run, bayesfast/utils/misc.py:62-108 (the uniqueness check is the caller's)"""
    nodes = np.asarray(nodes, dtype=np.float64)
    n_node = nodes.size
    w = np.ones(n_node - 1) / (n_node - 1) if weights is None else np.asarray(weights, dtype=np.float64) / np.sum(weights)
    a = np.asarray(a, dtype=np.float64)
    n_w = (n * w).astype(np.int64)
    n_w[-1] += n - np.sum(n_w)
    n_c = np.cumsum(np.insert(n_w, 0, 0))
    i_all = np.empty(n, dtype=np.int64)
    m = len(a)
    for j in range(n_node - 1):
        ep = (j == n_node - 2)
        i_j = np.linspace(nodes[j] * (m - 1) / 100, nodes[j + 1] * (m - 1) / 100, n_w[j], ep)
        i_all[n_c[j]:n_c[j + 1]] = i_j.astype(np.int64)
    return np.argsort(a, kind='stable')[i_all], i_all


def importance_weights(logp, logq, k_trunc):
    """PostStep, bayesfast/core/recipe.py:1286-1297"""
    weights = np.exp(np.asarray(logp) - np.asarray(logq))
    if k_trunc < 0:
        return weights, weights.copy()
    return weights, np.clip(weights, 0, np.mean(weights) * weights.size**k_trunc)
