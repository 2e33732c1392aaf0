"""ctypes bindings to oracle/_ref/libref_lobpcg.so — the UNMODIFIED reference library built by
oracle/Makefile plus oracle/ref_harness.c.  TEST INFRASTRUCTURE ONLY: imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by lobpcg_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "_ref" / "libref_lobpcg.so"

_P = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex64): "c", np.dtype(np.complex128): "z"}
_R = {"s": np.float32, "d": np.float64, "c": np.float32, "z": np.float64}


def available() -> bool:
    return LIB_PATH.exists()


_lib = None


def lib():
    global _lib
    if _lib is None:
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")  # OpenBLAS pthreads + OpenMP matvec
        _lib = C.CDLL(str(LIB_PATH))
        _lib.ref_blas_config.restype = C.c_char_p
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefOp:
    """Owns a LinearOperator_<p>_t* created by the harness; keeps the numpy buffers alive."""

    def __init__(self, prefix, handle, n, keep=()):
        self.prefix, self.handle, self.n, self._keep = prefix, handle, n, keep

    def apply(self, X):
        X = np.asfortranarray(X)
        Y = np.empty_like(X, order="F")
        f = getattr(lib(), f"ref_{self.prefix}_op_apply")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        f.restype = None
        f(self.handle, _ptr(X), _ptr(Y), X.shape[1] if X.ndim == 2 else 1)
        return Y

    def __del__(self):
        try:
            f = getattr(lib(), f"ref_{self.prefix}_op_free")
            f.argtypes = [C.c_void_p]
            f(self.handle)
        except Exception:
            pass


def op_stencil(grid, dtype, cdiag=None, coff=-1.0, potential=None):
    p = _P[np.dtype(dtype)]
    g = tuple(grid) + (1,) * (3 - len(grid))
    cdiag = 2.0 * len(grid) if cdiag is None else cdiag
    v = None if potential is None else np.ascontiguousarray(potential, dtype=_R[p])
    f = getattr(lib(), f"ref_{p}_op_stencil")
    f.argtypes = [C.c_int64] * 3 + [C.c_double] * 2 + [C.c_void_p]
    f.restype = C.c_void_p
    return RefOp(p, f(g[0], g[1], g[2], cdiag, coff, _ptr(v)), g[0] * g[1] * g[2], (v,))


def op_csr(rowptr, col, val):
    p = _P[val.dtype]
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val)
    f = getattr(lib(), f"ref_{p}_op_csr")
    f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    f.restype = C.c_void_p
    n = len(rowptr) - 1
    return RefOp(p, f(n, _ptr(rowptr), _ptr(col), _ptr(val)), n, (rowptr, col, val))


def op_diag(d, dtype):
    p = _P[np.dtype(dtype)]
    d = np.ascontiguousarray(d, dtype=_R[p])
    f = getattr(lib(), f"ref_{p}_op_diag")
    f.argtypes = [C.c_int64, C.c_void_p]
    f.restype = C.c_void_p
    return RefOp(p, f(len(d), _ptr(d)), len(d), (d,))


def op_bdg(grid, dtype, shift, d, cdiag=None, coff=-1.0):
    p = _P[np.dtype(dtype)]
    g = tuple(grid) + (1,) * (3 - len(grid))
    cdiag = 2.0 * len(grid) if cdiag is None else cdiag
    f = getattr(lib(), f"ref_{p}_op_bdg")
    f.argtypes = [C.c_int64] * 3 + [C.c_double] * 5
    f.restype = C.c_void_p
    d = complex(d)
    return RefOp(p, f(g[0], g[1], g[2], cdiag, coff, shift, d.real, d.imag), 2 * g[0] * g[1] * g[2])


def op_cheb(A: RefOp, degree, lo, hi):
    """T = p(A) for the reference solver: Chebyshev steps through A's own matvec (oracle/ref_harness.c)."""
    f = getattr(lib(), f"ref_{A.prefix}_op_cheb")
    f.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double]
    f.restype = C.c_void_p
    return RefOp(A.prefix, f(A.handle, int(degree), float(lo), float(hi)), A.n, (A,))


def solve(A, X0, nev, tol, max_iter, B=None, T=None, indefinite=False, verbosity=0):
    """Run <p>_lobpcg / <p>_ilobpcg of the reference.  Returns dict(eig,res,X,iter,converged,sig)."""
    X = np.array(X0, order="F", copy=True)
    n, k = X.shape
    p = _P[X.dtype]
    rt = _R[p]
    eig = np.zeros(k, dtype=rt)
    res = np.zeros(k, dtype=rt)
    sig = np.zeros(3 * k, dtype=np.int8)
    it = C.c_uint64(0)
    cv = C.c_uint64(0)
    f = getattr(lib(), f"ref_{p}_solve")
    f.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64),
                  C.POINTER(C.c_uint64), C.c_int]
    f.restype = C.c_int
    f(int(indefinite), n, nev, k, max_iter, float(tol), A.handle, B.handle if B else None,
      T.handle if T else None, _ptr(X), _ptr(eig), _ptr(res), _ptr(sig), C.byref(it), C.byref(cv), verbosity)
    return dict(eig=eig, res=res, X=X, iter=it.value, converged=cv.value, sig=sig)


def set_threads(n: int):
    lib().ref_set_threads(int(n))


def blas_config() -> str:
    return lib().ref_blas_config().decode()
