"""Warp-pair NUTS kernel (integrator warp + tree warp per 8-chain group, bfb_sampler_pair.cu) against the one-warp kernel:
bitwise comparison of every output on a small run, then rates at d=26 cubic-2.  usage: pair_probe.py [check|rate|both] [C,...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic
n = 26
mode = sys.argv[1] if len(sys.argv) > 1 else 'both'
Cs = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [4096, 16384]


def setup(C):
    prob = synthetic.des_shaped(n, seed=1, n_chain=C)
    sur = bfb.PolyModel('cubic-2', input_size=n, output_size=1)
    sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
    den = bfb.Density(sur)
    return prob, den._sync(False)


if mode in ('check', 'both'):
    for C, n_iter in ((8, 30), (70, 60), (1000, 120)):
        prob, h = setup(C)
        outs = {}
        for fam in ('dmma', 'pair'):
            os.environ['BFB200_SAMPLER'] = fam
            cfg = bfb.NTrace(n_chain=C, n_iter=n_iter, n_warmup=n_iter // 2, x_0=prob['x_0'])._cfg_dict(1, 0)
            h.sampler_init(cfg, prob['x_0'], 1. / n**0.25, np.ones(n), prob['x_0'])
            a = h.sampler_run('NUTS', n_iter // 3)
            b = h.sampler_run('NUTS', n_iter - n_iter // 3)
            outs[fam] = (a, b, h.sampler_state(), h.sampler_last_path())
        bad = 0
        for part in (0, 1):
            for k, v in outs['dmma'][part].items():
                w = outs['pair'][part][k]
                same = np.array_equal(np.asarray(v), np.asarray(w), equal_nan=True) if isinstance(v, np.ndarray) else v == w
                if not same:
                    bad += 1
                    d = np.asarray(v) != np.asarray(w)
                    print('DIFF', C, part, k, int(d.sum()), 'of', d.size, 'first at', np.argwhere(d)[:3].tolist())
        for k, v in outs['dmma'][2].items():
            if not np.array_equal(np.asarray(v), np.asarray(outs['pair'][2][k]), equal_nan=True):
                bad += 1
                print('DIFF state', C, k)
        print(json.dumps(dict(check=C, n_iter=n_iter, paths=[outs['dmma'][3], outs['pair'][3]], mismatching_fields=bad,
                              mean_tree=float(outs['dmma'][1]['tree_size'].mean()))), flush=True)

if mode in ('rate', 'both'):
    for C in Cs:
        prob, h = setup(C)
        peak = h.fp64_peak(0)
        for fam in ('pair', 'dmma'):
            os.environ['BFB200_SAMPLER'] = fam
            cfg = bfb.NTrace(n_chain=C, n_iter=1500, n_warmup=500, x_0=prob['x_0'])._cfg_dict(1, 0)
            h.sampler_init(cfg, prob['x_0'], 1. / n**0.25, np.ones(n), prob['x_0'])
            for k in (500, 500, 500):
                r = h.sampler_run('NUTS', k, fields=('tree_depth',))
                ms = h.last_kernel_ms()
                rate = r['total_tree_size'] / ms * 1e3
                print(json.dumps(dict(C=C, family=fam, iters=k, kernel=h.sampler_last_path(), ms=round(ms, 3), leapfrogs_per_s=rate,
                                      frac_fp64=round(rate * (8 * n * n + 24 * n) / 1e12 / peak, 4))), flush=True)
