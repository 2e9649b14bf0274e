"""BASELINE config 5: PolyModel cubic-2 fit sweep, d=32 (P=1585), N = 1e4 .. 1e7 rows resident on the device.
Reports the Gram kernel time (CUDA events on the launching stream) and algorithmic TFLOP/s = N P (P+1) / t."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bayesfast_b200 as bfb
from bayesfast_b200 import _cabi

n = int(os.environ.get('FIT_N', 32))
Ns = [int(float(v)) for v in (sys.argv[1:] or ['1e4', '1e5', '1e6', '1e7'])]
dev = torch.device('cuda:0')
sur = bfb.PolyModel('cubic-2', input_size=n, output_size=1)
P = sur.n_param
h = sur._dev()
h.set_model(sur.to_spec(with_bound=False))
L = _cabi.lib()
peak = h.fp64_peak(1)
g = torch.Generator(device=dev).manual_seed(0)
rows = []
for N in Ns:
    x = torch.randn(N, n, dtype=torch.float64, device=dev, generator=g)
    coef = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    y = (-0.5 * (x * x).sum(1) + x @ coef + 0.05 * (x ** 3).sum(1) + 1e-3 * torch.randn(N, dtype=torch.float64, device=dev, generator=g))[:, None].contiguous()
    torch.cuda.synchronize()
    best = 1e30
    for rep in range(3):
        _cabi.check(L.bfb_fit_begin(h._h, None))
        _cabi.check(L.bfb_fit_accumulate(h._h, x.data_ptr(), y.data_ptr(), None, N, _cabi.BFB_DEVICE))
        best = min(best, h.last_kernel_ms())
    t0 = time.time()
    out = np.empty(P)
    rr = C.c_double(0)
    _cabi.check(L.bfb_fit_solve(h._h, out.ctypes.data_as(_cabi._dp), C.byref(rr)))
    solve_s = time.time() - t0
    flops = float(N) * P * (P + 1) + 2. * N * P
    rows.append(dict(N=N, n=n, P=P, gram_ms=best, tflops=flops / (best * 1e-3) / 1e12, frac_of_dmma_peak=flops / (best * 1e-3) / 1e12 / peak,
                     solve_s=solve_s, rel_resid=rr.value, hbm_gbs=N * (n + 2) * 8 / (best * 1e-3) / 1e9))
    print(json.dumps(rows[-1]))
    del x, y
print(json.dumps(dict(dmma_peak_tflops=peak)))
