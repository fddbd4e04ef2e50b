"""Drop-in for the reference's ctypes layer SpaDOT/utils/OT_loss/ot_func.py, bound to libot_b200.so.

Same module-level functions, argument order and in-place semantics as the reference file
(`dummy_c`, `primal_c`, `dual_c`, `compute_duality_gap_c`, `update_K_c`, `update_R_c`,
`step1_process_c`, `update_process_c`; ot_func.py:317-567), so `ot_solvers.py` can import them
from here unchanged.  The library behind them is include/libot_b200.h: the reference's own
fourteen exports, executed by CUDA kernels on the B200.  There is no CPU fallback: importing
this module needs the built library, calling into it needs the device (the library aborts the
process otherwise, because libot's ABI has no error channel).

The streamed solver (`spadot_b200.ot_solvers`) is the fast path; this layer keeps K, _K, C and
R on the host between calls like the reference does and is therefore PCIe-bound.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
from numpy.ctypeslib import ndpointer

cur_folder = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(cur_folder, "libot_b200.so")
if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(spadot_b200 has no CPU fallback)")
lib = ctypes.cdll.LoadLibrary(LIB_PATH)          # ot_func.py:10


def _vec(t):
    return ndpointer(dtype=t, ndim=1, flags="C_CONTIGUOUS")


def _mat(t):
    return ndpointer(dtype=t, ndim=2, flags="C_CONTIGUOUS")


for _t, _sfx in ((ctypes.c_float, "float"), (ctypes.c_double, "double")):
    # (C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, m, n)                        ot_func.py:12-162
    for _name in ("dummy", "primal", "dual", "compute_duality_gap"):
        _fn = getattr(lib, f"{_name}_{_sfx}")
        _fn.argtypes = [_mat(_t)] * 3 + [_vec(_t)] * 6 + [_t] * 3 + [ctypes.c_int] * 2
        _fn.restype = _t
    _fn = getattr(lib, f"update_k_{_sfx}")                                                  # ot_func.py:164-188
    _fn.argtypes = [_mat(_t)] * 3 + [_vec(_t)] * 2 + [_t, ctypes.c_int, ctypes.c_int]
    _fn.restype = None
    _fn = getattr(lib, f"update_R_{_sfx}")                                                  # ot_func.py:190-210
    _fn.argtypes = [_mat(_t)] * 2 + [_vec(_t)] * 2 + [ctypes.c_int, ctypes.c_int]
    _fn.restype = None

_D, _I = ctypes.c_double, ctypes.c_int
lib.step1_process_double.argtypes = ([_vec(_D)] * 4 + [_mat(_D)] * 2 + [_vec(_D)] * 6 + [_I] * 3 + [_D] * 6 + [_I] * 2)
lib.step1_process_double.restype = _I                                                       # ot_func.py:258-284
# all 28 parameters are declared here (the reference declares 26 and relies on ctypes' default int conversion for
# the trailing m, n; ot_func.py:286-313 vs :561-567)
lib.update_process_double.argtypes = ([_mat(_D)] + [_vec(_D)] * 4 + [_mat(_D)] * 3 + [_vec(_D)] * 6 + [_I] * 3 + [_D] * 7
                                      + [_I] * 4)
lib.update_process_double.restype = _D
lib.libot_b200_device_check.restype = _I
lib.libot_b200_version.restype = _I
lib.libot_b200_counters.argtypes = [ctypes.POINTER(ctypes.c_longlong)] * 3
lib.libot_b200_counters.restype = None
lib.libot_b200_release.argtypes = []
lib.libot_b200_release.restype = None


def counters():
    """(kernel launches, bytes uploaded, bytes downloaded) since the library was loaded."""
    v = [ctypes.c_longlong(0) for _ in range(3)]
    lib.libot_b200_counters(*[ctypes.byref(x) for x in v])
    return tuple(int(x.value) for x in v)


def release():
    """Return the library's pooled device memory to the driver."""
    lib.libot_b200_release()


def _gap_like(name, C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, use_float):
    m, n = C.shape
    t, sfx = (ctypes.c_float, "float") if use_float else (ctypes.c_double, "double")
    arr = [np.ascontiguousarray(x, dtype=t) for x in (C, K, R, dx, dy, p, q, a, b)]
    return getattr(lib, f"{name}_{sfx}")(*arr, epsilon, lambda1, lambda2, m, n)


def dummy_c(C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, use_float=False):          # ot_func.py:317
    return _gap_like("dummy", C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, use_float)


def primal_c(C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, use_float=False):         # ot_func.py:345
    return _gap_like("primal", C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, use_float)


def dual_c(C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, use_float=False):           # ot_func.py:373
    return _gap_like("dual", C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, use_float)


def compute_duality_gap_c(C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, use_float=False):   # ot_func.py:401
    return _gap_like("compute_duality_gap", C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, use_float)


def update_K_c(K, _K, C, u, v, epsilon, use_float=False):                                      # ot_func.py:429
    m, n = C.shape
    (lib.update_k_float if use_float else lib.update_k_double)(K, _K, C, u, v, epsilon, m, n)


def update_R_c(R, K, a, b, use_float=False):                                                   # ot_func.py:456
    m, n = K.shape
    (lib.update_R_float if use_float else lib.update_R_double)(R, K, a, b, m, n)


def step1_process_c(a, b, old_a, old_b, K, C, dx, dy, p, q, u, v, cur_iter, max_iter, iters, tau, lambda1, lambda2,
                    alpha1, alpha2, epsilon):                                                   # ot_func.py:521
    m, n = K.shape
    return lib.step1_process_double(a, b, old_a, old_b, K, C, dx, dy, p, q, u, v, cur_iter, int(max_iter), iters, float(tau),
                                    lambda1, lambda2, alpha1, alpha2, epsilon, m, n)


def update_process_c(R, a, b, old_a, old_b, K, _K, C, dx, dy, p, q, u, v, epsilon_scaling, cur_epsilon_scaling, batch_size,
                     epsilon, threshold, tau, lambda1, lambda2, alpha1, alpha2, cur_iter, max_iter):   # ot_func.py:552
    m, n = K.shape
    return lib.update_process_double(R, a, b, old_a, old_b, K, _K, C, dx, dy, p, q, u, v, epsilon_scaling,
                                     cur_epsilon_scaling, batch_size, epsilon, threshold, float(tau), lambda1, lambda2,
                                     alpha1, alpha2, cur_iter, int(max_iter), m, n)
