"""GPU parity of the tempered samplers TNUTS / THMC (csrc/bfb_sampler_tempered.cu) vs recorded runs of the real reference
(tests/golden/sampler_tempered.npz) and vs the oracle (samplers/tnuts.py, thmc.py, hmc_utils/base_hmc.py:220-262,
hmc_utils/integration.py:98-222)."""
import numpy as np
import pytest

import _golden_io as gio
from _specs import to_device_spec, synthetic_spec
from test_gpu_sampler import cfg_from, device_draws, check_floats, INT_STATS, FLT_STATS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def handles():
    from bayesfast_b200 import _cabi
    h, hb = _cabi.Handle(0), _cabi.Handle(0)
    yield h, hb
    h.close()
    hb.close()


def compare(out, r, sampler, late=1e-3):
    if sampler == 'TNUTS':
        for k in INT_STATS:
            assert np.array_equal(out[k], np.asarray(r[k]).astype(np.int32)), k
        for k in FLT_STATS + ('u', 'weight'):
            check_floats(out[k], r[k], k, late)
    else:
        assert np.array_equal(out['tree_depth'], np.asarray(r['accepted' if 'accepted' in r else 'tree_depth']).astype(np.int32))
        assert np.array_equal(out['diverging'], np.asarray(r['diverging']).astype(np.int32))
        for k, k2 in (('logp', 'logp'), ('energy', 'energy'), ('mean_tree_accept', 'accept_stat'), ('step_size', 'step_size'),
                      ('energy_change', 'energy_change'), ('u', 'u'), ('weight', 'weight')):
            check_floats(out[k], r[k2 if k2 in r else k], k, late)
    check_floats(out['samples'], r['samples'], 'samples', late)


@pytest.mark.parametrize('case', gio.load('sampler_tempered.npz')['cases'], ids=lambda c: c['name'])
def test_tempered_golden_chains(handles, case):
    """same seeds, draw stream and u_0 as the recorded TNUTS / THMC runs of the real reference"""
    h, hb = handles
    r, kw = case['result'], case['trace_kw']
    n_iter, n_warmup = int(kw['n_iter']), int(kw['n_warmup'])
    h.set_model(to_device_spec(case['spec']))
    hb.set_model(to_device_spec(case['base_spec']))
    cfg = cfg_from(kw, n_warmup, int(case['seed']))
    h.tsampler_init(hb, float(case['logxi']), cfg, case['x0'], case['u0'], float(r['step0']), r['var0'], case['x0'])
    out = h.tsampler_run(case['sampler'], n_iter)
    st = h.sampler_state()
    assert np.all(st['status'] == 0)
    assert np.array_equal(st['n_draws'], r['n_draws'])
    assert h.sampler_last_path() == 'generic'
    # the reference's THMC lets u run away (|u| ~ 500 in the recorded run): late-iteration floats are compared loosely there
    compare(out, r, case['sampler'], late=1e-3 if case['sampler'] == 'TNUTS' else 5e-3)
    if case['sampler'] == 'TNUTS':
        assert out['total_tree_size'] == int(r['tree_size'].sum())
    assert np.allclose(st['final_step'], r['final_step'], rtol=5e-3)
    assert np.allclose(st['final_var'], r['final_var'], rtol=5e-3)


@pytest.mark.parametrize('sampler,n,order,C,n_iter', [('TNUTS', 26, 'cubic-2', 96, 30), ('TNUTS', 40, 'cubic-3', 8, 12),
                                                      ('THMC', 26, 'cubic-2', 64, 20), ('TNUTS', 2, 'quadratic', 40, 40)])
def test_tempered_teacher_forced_vs_oracle(handles, oracle, sampler, n, order, C, n_iter):
    """integer outcomes (tree depths / sizes / divergences, acceptances) and draw counts identical to the oracle fed with the
    device's own draws; u, weight and the samples within the float windows; chunked runs and reset are bit-identical"""
    h, hb = handles
    spec, cov = synthetic_spec(n, order, seed=11 + n, decay=True)
    bspec, _ = synthetic_spec(n, 'quadratic', seed=3 + n, cond=4., bound=False)
    for cf in bspec['configs']:                                  # a broad base density: the quadratic of the target's shape / 3
        cf['coef'] = np.asarray(cf['coef']) / 3.
    h.set_model(to_device_spec(spec))
    hb.set_model(to_device_spec(bspec))
    rng = np.random.default_rng(5)
    x0 = (np.linalg.cholesky(cov) @ rng.normal(size=(n, C))).T
    u0 = rng.normal(size=C)
    seed, chain0, logxi = 777, 300, 0.25
    kw = dict(n_int_step=5) if sampler == 'THMC' else {}
    cfg = cfg_from(kw, n_iter // 2, seed, chain0)
    step0 = (0.3 if sampler == 'THMC' else 1.) / n**0.25
    h.tsampler_init(hb, logxi, cfg, x0, u0, step0, np.ones(n), x0)
    out = h.tsampler_run(sampler, n_iter)
    st = h.sampler_state()
    U, Z = device_draws(h, seed, st['n_draws'], chain0)
    ocfg = dict(n_iter=n_iter, n_warmup=n_iter // 2, **kw)
    ref = oracle.OracleDensity(spec).run(sampler, ocfg, x0, step0, np.ones(n), draws_u=U, draws_z=Z,
                                         base=oracle.OracleDensity(bspec), logxi=logxi, u0=u0)
    # THMC records u of the integrated state even when the step diverged (thmc.py:18), so a divergence with a non-finite u ends
    # the chain with 'Bad initial energy' (base_hmc.py:249-253) in the reference, the oracle and here alike: same chains, same status
    assert np.array_equal(st['status'], ref['status'])
    ok = st['status'] == 0
    assert np.all(ok) if sampler == 'TNUTS' else np.sum(ok) >= C // 2
    assert np.array_equal(st['n_draws'], ref['n_draws'])
    if sampler == 'TNUTS':
        ref_r = ref
        assert int(np.max(ref['tree_depth'])) >= 3
    else:
        ref_r = dict(ref, accepted=ref['tree_depth'], accept_stat=ref['mean_tree_accept'])
    compare({k: v[ok] for k, v in out.items() if isinstance(v, np.ndarray) and v.shape[:1] == (C,)},
            {k: v[ok] for k, v in ref_r.items() if isinstance(v, np.ndarray) and v.shape[:1] == (C,)}, sampler)
    # two calls of half the length and a reset reproduce the single call bit for bit
    h.sampler_reset()
    a = h.tsampler_run(sampler, n_iter // 2)
    b = h.tsampler_run(sampler, n_iter - n_iter // 2)
    for k in ('samples', 'u', 'weight', 'tree_size', 'energy'):
        assert np.array_equal(np.concatenate((a[k], b[k]), axis=1)[ok], out[k][ok]), k


def test_tempered_sample_api():
    """bayesfast_b200.sample with a TNTrace / THTrace: TraceTuple of TNTrace objects whose stats carry 'u' and 'weight'
    (hmc_utils/stats.py:9-24), get(return_type='weights') (sample_trace.py:574-585)"""
    import bayesfast_b200 as bfb
    from _specs import pack
    n = 6

    def density(order, seed, scale=1.):
        spec, cov = synthetic_spec(n, order, seed=seed, bound=False)
        sur = bfb.PolyModel(order, input_size=n, output_size=1)
        for conf, cf in zip(sur.configs, spec['configs']):
            conf._set(pack(cf['order'], np.asarray(cf['coef'][0]) * scale, n), 0)
        return bfb.Density(sur, decay_options=dict(use_decay=False)), cov

    den, cov = density('cubic-2', 4)
    base, _ = density('quadratic', 4, 0.4)
    x0 = (np.linalg.cholesky(cov) @ np.random.default_rng(1).normal(size=(n, 12))).T * 0.5
    u0 = np.linspace(-1., 1., 12)
    tr = bfb.TNTrace(base, 0.1, n_chain=12, n_iter=40, n_warmup=20, x_0=x0, random_generator=5, u_0=u0)
    tt = bfb.sample(den, tr, verbose=False)
    assert tt.sampler == 'TNUTS' and tt.samples.shape == (12, 40, n)
    t3 = tt[3]
    assert isinstance(t3, bfb.TNTrace) and t3.chain_id == 3
    assert t3.stats.stats_items[:2] == ('u', 'weight') and len(t3.stats._u) == 40
    w = t3.get(return_type='weights')
    assert w.shape == (20,) and np.all(w > 0) and np.all(np.isfinite(w))
    assert tt.n_call == int(np.sum(tt.arrays['tree_size'][:, 1:])) + 12 * 41
    # same seeds -> same chains; a different u_0 -> different chains
    tt2 = bfb.sample(den, bfb.TNTrace(base, 0.1, n_chain=12, n_iter=40, n_warmup=20, x_0=x0, random_generator=5, u_0=u0),
                     verbose=False)
    assert np.array_equal(tt2.samples, tt.samples) and np.array_equal(tt2.arrays['u'], tt.arrays['u'])
    tt3 = bfb.sample(den, bfb.TNTrace(base, 0.1, n_chain=12, n_iter=40, n_warmup=20, x_0=x0, random_generator=5, u_0=u0 + 0.5),
                     verbose=False)
    assert not np.array_equal(tt3.samples, tt.samples)
    # THMC: the reference's sampler ends a chain with 'Bad initial energy' once a divergent step left a non-finite u behind
    # (thmc.py:18, base_hmc.py:249-253); a short, fixed step keeps this run clear of that
    try:
        th = bfb.sample(den, bfb.THTrace(base, 0.1, n_chain=12, n_iter=30, n_warmup=15, n_int_step=3, x_0=x0, random_generator=5,
                                         u_0=u0, step_size=0.15, adapt_step_size=False), verbose=False)
        assert th.sampler == 'THMC' and isinstance(th[0], bfb.THTrace) and th[0].stats._n_int_step[0] == 3
        assert th[0].stats.stats_items[:2] == ('u', 'weight') and th[0].get(return_type='weights').shape == (15,)
    except RuntimeError as e:
        assert 'Bad initial energy' in str(e)
    with pytest.raises(NotImplementedError):
        bfb.sample(den, bfb.TNTrace(base, n_chain=4, n_iter=10, n_warmup=5, x_0=x0[:4], metric='full'), verbose=False)
    with pytest.raises(ValueError):
        bfb.sample(den, bfb.TNTrace(den, n_chain=4, n_iter=10, n_warmup=5, x_0=x0[:4]), verbose=False)
