"""Where the warm time of PolyModel.fit goes at the headline size (n=26 cubic-2, N=4216, P=1054): cProfile of the host path."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bayesfast_b200 as bfb
from bayesfast_b200 import synthetic
n = 26
prob = synthetic.des_shaped(n, seed=1, n_chain=64)
sur = bfb.PolyModel('cubic-2', input_size=n, output_size=1)
for i in range(3):
    t0 = time.time()
    sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
    print('fit', i, round((time.time() - t0) * 1e3, 2), 'ms')
pr = cProfile.Profile()
pr.enable()
for i in range(5):
    sur.fit(prob['x_fit'], prob['y_fit'], logp=prob['y_fit'][:, 0])
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
