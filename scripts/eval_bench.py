"""Throughput of the batched surrogate evaluation (bfb_logp_and_grad_batch, device-resident points) against the FP64 peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
from bayesfast_b200 import _cabi
from _specs import synthetic_spec, to_device_spec

h = _cabi.Handle(0)
peak = max(h.fp64_peak(0) for _ in range(2))
for n, order, C in ((26, 'cubic-2', 1 << 22), (26, 'cubic-2', 4096), (26, 'cubic-2', 32768), (16, 'cubic-2', 1 << 22), (32, 'cubic-2', 1 << 22), (26, 'quadratic', 1 << 22)):
    spec, cov = synthetic_spec(n, order, seed=1)
    h.set_model(to_device_spec(spec))
    L = torch.tensor(np.linalg.cholesky(cov), device='cuda')
    X = (torch.randn(C, n, dtype=torch.float64, device='cuda') @ L.T).contiguous()
    lp = torch.empty(C, dtype=torch.float64, device='cuda'); g = torch.empty(C, n, dtype=torch.float64, device='cuda')
    torch.cuda.synchronize()
    flops = (8 * n * n + 15 * n if order == 'cubic-2' else 4 * n * n + 9 * n) * C
    for mode in ('dmma', 'generic'):
        os.environ['BFB200_EVAL'] = mode
        ms = []
        for _ in range(5):
            h.logp_and_grad_batch_dev(X.data_ptr(), C, lp.data_ptr(), g.data_ptr())
            ms.append(h.last_kernel_ms())
        best = min(ms[1:])
        print(json.dumps(dict(n=n, order=order, C=C, mode=mode, ms=round(best, 4), points_per_s=C / best * 1e3,
                              tflops=flops / best / 1e9, frac_fp64=flops / best / 1e9 / peak,
                              gbs=C * (2 * n + 1) * 8 / best / 1e6, peak_tflops=peak)))
