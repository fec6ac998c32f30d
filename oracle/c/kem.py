"""ctypes wrapper of the oracle's C port (oracle/c/kalman_em.c).  Oracle infrastructure."""
import ctypes as C
import numpy as np
from . import build_c

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_c.build())
        _lib.kem_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5 + \
                                  [C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.kem_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return int(lib().kem_max_threads())


def em_kalman_batch(X, Lam, R, A, Q, p=1, P0=None, max_iter=50, tol=0.0, nthreads=0, want_F=True):
    """X (B,T,N); Lam (B,N,r); R (B,N); A (B,r,k); Q (B,r,r).  Parameters are updated on copies."""
    X = np.ascontiguousarray(X, float); B, T, N = X.shape; r = Lam.shape[-1]
    Lam = np.array(Lam, float, order="C"); R = np.array(R, float, order="C")
    A = np.array(A, float, order="C"); Q = np.array(Q, float, order="C")
    P0c = np.ascontiguousarray(P0, float) if P0 is not None else None
    F = np.empty((B, T, r)) if want_F else None
    ll = np.empty((B, max_iter)); it = np.zeros(B, np.int32); st = np.zeros(B, np.int32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    lib().kem_batch(ptr(X), B, T, N, r, p, ptr(Lam), ptr(R), ptr(A), ptr(Q), ptr(P0c), max_iter, tol, ptr(F), ptr(ll), ptr(it),
                    ptr(st), nthreads)
    return dict(Lam=Lam, R=R, A=A, Q=Q, F=F, loglik=ll, iters=it, status=st)
