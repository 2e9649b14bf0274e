import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesfast_b200 import _cabi
h = _cabi.Handle(0)
L = _cabi.lib()
L.bfb_dmma_issue_test.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
for wps in (4, 8, 16):
    for nacc, src in ((1, 0), (2, 0), (4, 0), (8, 0), (15, 0), (8, 1), (15, 1)):
        v = C.c_double()
        rc = L.bfb_dmma_issue_test(h._h, nacc, src, wps, C.byref(v))
        print('warps/SM', wps, 'nacc', nacc, 'src', 'smem' if src else 'reg', 'cycles per DMMA (per warp) %.1f' % v.value, 'rc', rc)
