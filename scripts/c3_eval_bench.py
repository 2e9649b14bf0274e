"""Generic evaluator at n = 64: what the cubic-3 terms cost (config[3] of BASELINE.json)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from bayesfast_b200 import _cabi
from _specs import synthetic_spec, to_device_spec
h = _cabi.Handle(0)
n, C = 64, 1 << 16
for order in ('cubic-2', 'cubic-3'):
    spec, cov = synthetic_spec(n, order, seed=3, bound=False)
    h.set_model(to_device_spec(spec))
    X = torch.randn(C, n, dtype=torch.float64, device='cuda').contiguous()
    lp = torch.empty(C, dtype=torch.float64, device='cuda'); g = torch.empty(C, n, dtype=torch.float64, device='cuda')
    ms = []
    for _ in range(4):
        h.logp_and_grad_batch_dev(X.data_ptr(), C, lp.data_ptr(), g.data_ptr()); ms.append(h.last_kernel_ms())
    best = min(ms[1:])
    fl = 8 * n * n + 15 * n + (1.5 * n * (n - 1) * (n - 2) if order == 'cubic-3' else 0)
    print(json.dumps(dict(order=order, n=n, C=C, ms=round(best, 3), points_per_s=C / best * 1e3, tflops=fl * C / best / 1e9)))
