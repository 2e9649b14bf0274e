"""
Pool of page-locked host buffers for sampler outputs.  cudaHostAlloc is slow (page pinning), so buffers are
recycled: a numpy array handed out by `empty()` returns its buffer to the pool when it is garbage collected.
"""
import ctypes as C
import threading
import weakref

import numpy as np

from . import _cabi

_free = {}          # nbytes -> [ptr, ...]
_lock = threading.RLock()      # re-entrant: _release runs in GC finalizers, which may fire while this thread holds the lock
_MAX_POOLED = 24 << 30
_pooled = 0                    # bytes currently in the pool (kept as a counter: no allocation under the lock)


def _release(ptr, nbytes):
    global _pooled
    with _lock:
        if _pooled + nbytes <= _MAX_POOLED:
            _free.setdefault(nbytes, []).append(ptr)
            _pooled += nbytes
            return
    _cabi.lib().bfb_host_free(C.c_void_p(ptr))


def empty(shape, dtype=np.float64):
    """uninitialised pinned numpy array"""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    nbytes_al = max(4096, (nbytes + 4095) // 4096 * 4096)
    global _pooled
    ptr = None
    with _lock:
        lst = _free.get(nbytes_al)
        if lst:
            ptr = lst.pop()
            _pooled -= nbytes_al
    if ptr is None:
        p = C.c_void_p()
        L = _cabi.lib()
        L.bfb_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        L.bfb_host_free.argtypes = [C.c_void_p]
        _cabi.check(L.bfb_host_alloc(nbytes_al, C.byref(p)))
        ptr = p.value
    buf = (C.c_char * nbytes_al).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, _release, ptr, nbytes_al)
    return arr


def trim():
    """free every pooled buffer"""
    global _pooled
    with _lock:
        items = [(k, p) for k, v in _free.items() for p in v]
        _free.clear()
        _pooled = 0
    for _, p in items:
        _cabi.lib().bfb_host_free(C.c_void_p(p))
