"""ctypes driver of oracle/_ref/libot_ref.so.  TEST INFRASTRUCTURE ONLY.

libot_ref.so is the reference's own native file (SpaDOT/utils/OT_loss/ot_func.cpp)
compiled UNMODIFIED, in place, by oracle/Makefile (`make ref`).  This module only
restates the thin Python driver around it (ot_solvers.py:240-290, the use_C and
c_for_v2 branch) so the native reference can run on the GPU box where
/root/reference does not exist.  It is the `"kind": "reference"` CPU baseline.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libot_ref.so")

_D = ctypes.c_double
_I = ctypes.c_int
_P = ctypes.c_void_p
_lib = None


def available() -> bool:
    return os.path.exists(REF_SO)


def bind(path):
    """ctypes binding of a library with libot's ABI (the compiled reference, or any drop-in for it)."""
    L = ctypes.CDLL(path)
    # ot_func.cpp:1203 update_k_double(K,K_,C,u,v,eps,m,n)
    L.update_k_double.argtypes = [_P] * 5 + [_D, _I, _I]
    L.update_k_double.restype = None
    # ot_func.cpp:1243 update_R_double(R,K,a,b,m,n)
    L.update_R_double.argtypes = [_P] * 4 + [_I, _I]
    L.update_R_double.restype = None
    # ot_func.cpp:1261 step1_process_double(a,b,old_a,old_b,K,C,dx,dy,p,q,u,v,cur,max,iters,tau,l1,l2,a1,a2,eps,m,n)
    L.step1_process_double.argtypes = [_P] * 12 + [_I, _I, _I] + [_D] * 6 + [_I, _I]
    L.step1_process_double.restype = _I
    # ot_func.cpp:1313 update_process_double (28 parameters)
    L.update_process_double.argtypes = [_P] * 14 + [_I, _I, _I] + [_D] * 7 + [_I, _I, _I, _I]
    L.update_process_double.restype = _D
    # ot_func.cpp:1011,1079,1147 {primal,dual,compute_duality_gap}_double(C,K,R,dx,dy,p,q,a,b,eps,l1,l2,m,n)
    for name in ("primal_double", "dual_double", "compute_duality_gap_double"):
        fn = getattr(L, name)
        fn.argtypes = [_P] * 9 + [_D] * 3 + [_I, _I]
        fn.restype = _D
    return L


def lib():
    global _lib
    if _lib is None:
        _lib = bind(REF_SO)
    return _lib


def _ptr(x: np.ndarray):
    assert x.dtype == np.float64 and x.flags["C_CONTIGUOUS"]
    return x.ctypes.data_as(_P)


def step1(a, b, old_a, old_b, K, C, dx, dy, p, q, u, v, cur_iter, max_iter, iters, tau,
          lambda1, lambda2, alpha1, alpha2, eps, L=None):
    """One call of step1_process_double (ot_func.py:521-550): `iters` Sinkhorn updates in place."""
    m, n = K.shape
    return (L or lib()).step1_process_double(_ptr(a), _ptr(b), _ptr(old_a), _ptr(old_b), _ptr(K), _ptr(C),
                                      _ptr(dx), _ptr(dy), _ptr(p), _ptr(q), _ptr(u), _ptr(v),
                                      int(cur_iter), int(max_iter), int(iters), float(tau),
                                      float(lambda1), float(lambda2), float(alpha1), float(alpha2),
                                      float(eps), m, n)


def gap_parts(C, K_, R, dx, dy, p, q, a, b, eps, lambda1, lambda2, L=None):
    """(primal, dual, gap) through the three standalone exports (ot_func.py:345-427)."""
    L = L or lib()
    m, n = C.shape
    args = [_ptr(x) for x in (C, K_, R, dx, dy, p, q, a, b)] + [float(eps), float(lambda1), float(lambda2), m, n]
    return L.primal_double(*args), L.dual_double(*args), L.compute_duality_gap_double(*args)


def duality_gap_solve(C, G, lambda1, lambda2, epsilon, batch_size=5, tolerance=1e-8, tau=1000.0,
                      epsilon0=1.0, max_iter=1e7, info=None, L=None, **ignored):
    """optimal_transport_duality_gap with use_C=True, c_for_v2=True (ot_solvers.py:264-290):
    python does the stage bookkeeping, update_k_double + update_process_double do the rest."""
    L = L or lib()
    C = np.ascontiguousarray(C, dtype=np.float64)
    I, J = C.shape
    scale_factor = np.exp(-np.log(epsilon) / 5)
    dx, dy = np.ones(I) / I, np.ones(J) / J
    p = np.ascontiguousarray(G, dtype=np.float64)
    q = np.ones(J) * np.average(G)
    u, v = np.zeros(I), np.zeros(J)
    a, b = np.ones(I), np.ones(J)
    eps_i = epsilon0 * scale_factor
    K = np.empty_like(C)
    K_ = np.empty_like(C)
    R = np.zeros_like(C)
    gap = np.inf
    for e in range(6):
        u = np.ascontiguousarray(u + eps_i * np.log(a))
        v = np.ascontiguousarray(v + eps_i * np.log(b))
        a, b = np.ones(I), np.ones(J)
        eps_i = eps_i / scale_factor
        alpha1 = lambda1 / (lambda1 + eps_i)
        alpha2 = lambda2 / (lambda2 + eps_i)
        old_a, old_b = a.copy(), b.copy()
        threshold = tolerance if e == 5 else 1e-6
        L.update_k_double(_ptr(K), _ptr(K_), _ptr(C), _ptr(u), _ptr(v), float(eps_i), I, J)
        gap = L.update_process_double(
            _ptr(R), _ptr(a), _ptr(b), _ptr(old_a), _ptr(old_b), _ptr(K), _ptr(K_), _ptr(C),
            _ptr(dx), _ptr(dy), _ptr(p), _ptr(q), _ptr(u), _ptr(v),
            5, e, int(batch_size), float(eps_i), float(threshold), float(tau),
            float(lambda1), float(lambda2), float(alpha1), float(alpha2), 0, int(max_iter), I, J)
    if np.isnan(gap):
        raise RuntimeError("Overflow encountered in duality gap computation, please report this incident")
    if info is not None:
        info.update(gap=float(gap), f=u + eps_i * np.log(a), g=v + eps_i * np.log(b), epsilon_final=eps_i,
                    a=a, b=b, u=u, v=v, K=K)
    return R / J
