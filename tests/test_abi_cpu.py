"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/bfb200.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from _specs import gpu_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from bayesfast_b200 import _cabi
    assert os.path.exists(_cabi.LIB_PATH), 'build with `make -C bayesfast_b200/csrc`'
    L = _cabi.lib()
    declared = _cabi.exported_symbols()
    assert len(declared) >= 27
    for name in declared:
        assert hasattr(L, name), name
    dyn = subprocess.check_output(['nm', '-D', '--defined-only', _cabi.LIB_PATH]).decode()
    for name in declared:
        assert (' T ' + name) in dyn, name
    assert L.bfb_version() >= 100


def test_no_cpu_fallback():
    from bayesfast_b200 import _cabi
    if gpu_available():
        pytest.skip('a GPU is present')
    with pytest.raises(_cabi.BfbError, match='no CUDA device|no CPU fallback'):
        _cabi.Handle(0)
    import bayesfast_b200 as bfb
    s = bfb.PolyModel('quadratic', input_size=2, output_size=1)
    for c in s.configs:
        c._set(np.ones(c._a_shape), 0)
    with pytest.raises(_cabi.BfbError):
        s._fun(np.zeros(2))
    with pytest.raises(_cabi.BfbError):
        s.fit(np.random.default_rng(0).normal(size=(20, 2)), np.zeros((20, 1)))


def test_struct_layouts_match_header():
    """field order / sizes of the ctypes structures against the C header (compiled with gcc)"""
    from bayesfast_b200 import _cabi
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "bfb200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(bfb_model_desc), offsetof(bfb_model_desc, alpha), offsetof(bfb_model_desc, hard_bounds),
         sizeof(bfb_sampler_cfg), offsetof(bfb_sampler_cfg, chain0), sizeof(bfb_run_out));
  return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, 't.c'), 'w').write(src)
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), '-o', os.path.join(d, 't'), os.path.join(d, 't.c')])
        vals = [int(v) for v in subprocess.check_output([os.path.join(d, 't')]).split()]
    assert vals == [C.sizeof(_cabi.ModelDesc), _cabi.ModelDesc.alpha.offset, _cabi.ModelDesc.hard_bounds.offset,
                    C.sizeof(_cabi.SamplerCfg), _cabi.SamplerCfg.chain0.offset, C.sizeof(_cabi.RunOut)]
