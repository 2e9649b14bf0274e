"""
Times the UNMODIFIED reference (h3jia/bayesfast, vendored by __graft_entry__.build() into baseline/_ref, git-ignored) on the
host cores of this box: its own multiprocess chain pool -- bf.sample(density, sample_trace, sampler='NUTS') with
bf.utils.parallel.set_backend(nproc) (core/sample.py:185-214, utils/parallel.py:130-150) -- on the same synthetic posterior,
training set, x_0, n_iter / n_warmup as the GPU run of bench.py.  Runs as a subprocess of bench.py (the import shim below
must not leak into the benchmark process); prints one JSON line.

usage: python baseline/ref_runner.py inputs.npz nproc chains_per_proc
"""
import json
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    # numpy-2 aliases the reference still uses (poly.py:60, module.py:72, ...) and stubs of two plotting / numerical-derivative
    # packages that nothing on this path touches
    np.int = int
    np.float = float
    for name in ('matplotlib', 'matplotlib.pyplot', 'numdifftools'):
        sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    for a in ('Gradient', 'Hessian', 'Jacobian', 'Hessdiag'):
        setattr(sys.modules['numdifftools'], a, None)
    sys.path.insert(0, os.path.join(HERE, '_ref'))
    import bayesfast
    return bayesfast


def main():
    inp = np.load(sys.argv[1])
    nproc, per = int(sys.argv[2]), int(sys.argv[3])
    n_iter, n_warmup, order = int(inp['n_iter']), int(inp['n_warmup']), str(inp['order'])
    bf = import_reference()
    x_fit, y_fit, x_0 = inp['x_fit'], inp['y_fit'], inp['x_0']
    n = x_fit.shape[1]
    sur = bf.modules.PolyModel(order, input_size=n, output_size=1, input_vars='x', output_vars='logp')
    mod = bf.Module(fun=lambda x: np.zeros(1), input_vars='x', output_vars='logp')        # never called: use_surrogate below
    den = bf.Density(density_name='logp', module_list=[mod], surrogate_list=[sur], input_vars='x', input_shapes=n)
    t0 = time.time()
    sur.fit(x_fit, y_fit, logp=y_fit[:, 0])
    fit_s = time.time() - t0
    den.use_surrogate = True
    bf.utils.parallel.set_backend(nproc)
    n_chain = nproc * per
    t0 = time.time()
    tt = bf.sample(den, sample_trace=dict(n_chain=n_chain, n_iter=n_iter, n_warmup=n_warmup, x_0=x_0[:n_chain]), sampler='NUTS',
                   verbose=False)
    wall = time.time() - t0
    leaves = int(sum(int(np.sum(t.stats._tree_size)) for t in tt))
    ndiv = int(sum(int(np.sum(t.stats._diverging)) for t in tt))
    print(json.dumps(dict(leaves=leaves, wall_s=wall, value=leaves / wall, nproc=nproc, n_chain=n_chain, n_iter=n_iter,
                          n_warmup=n_warmup, fit_s=fit_s, n_diverging=ndiv,
                          mean_tree_size=leaves / float(n_chain * n_iter))))


if __name__ == '__main__':
    main()
