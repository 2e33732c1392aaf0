"""ctypes binding of liblobpcg_b200.so — the host-side mirror of the reference's solver interface.

Everything here goes through the C ABI declared in include/lobpcg_b200.h / include/lobpcg.h; there is no
Python or CPU compute path.  Importing works without a GPU (symbols can be inspected), but any call that
needs a device fails loudly.

Reference interface being mirrored: ``<p>_lobpcg_t`` / ``lobpcg(alg)`` / ``ilobpcg(alg)``
(reference lobpcg.h:10-92, 590-686) and ``LinearOperator_<p>_t`` (include/lobpcg/linop.h:7-26).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "_lib" / "liblobpcg_b200.so"

PREFIX = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex64): "c",
          np.dtype(np.complex128): "z"}
DTYPE = {v: k for k, v in PREFIX.items()}
REAL = {"s": np.dtype(np.float32), "d": np.dtype(np.float64), "c": np.dtype(np.float32), "z": np.dtype(np.float64)}


class LobpcgB200Error(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; no fallback of any kind exists."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise LobpcgB200Error(
                f"{LIB_PATH} is missing: build it with `python -m lobpcg_b200.build` (needs nvcc); "
                "lobpcg_b200 has no CPU fallback")
        _lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
        _declare(_lib)
    return _lib


def _declare(L):
    vp, i64, i32, u64, dbl, ci = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_double, C.c_int
    L.lb2_version.restype = C.c_char_p
    L.lb2_ctx_create.restype = vp
    L.lb2_ctx_create.argtypes = [ci, vp]
    L.lb2_ctx_destroy.argtypes = [vp]
    L.lb2_ctx_sync.argtypes = [vp]
    L.lb2_ctx_trim.argtypes = [vp]
    L.lb2_ctx_set_option.argtypes = [vp, C.c_char_p, ci]
    L.lb2_ctx_launches.restype = C.c_ulonglong
    L.lb2_ctx_launches.argtypes = [vp]
    L.lb2_ctx_oz_stats.argtypes = [vp, C.POINTER(dbl)]
    L.lb2_oz_plan_check.argtypes = [ci, ci, ci, ci, i64, ci, ci, C.POINTER(dbl)]
    L.lb2_default_ctx.restype = vp
    L.lb2_gram_wl_plan_check.argtypes = [ci, ci, ci, i64, ci, ci, C.POINTER(dbl)]
    L.lb2_gram_wl_cols_plan_check.argtypes = [ci, ci, ci, ci, i64, ci, ci, C.POINTER(dbl)]
    L.lb2_gram_wl_plan_sharing.argtypes = [ci, ci, ci, i64, ci, ci, ci, ci, ci, C.POINTER(dbl)]
    L.lb2_malloc.restype = vp
    L.lb2_malloc.argtypes = [C.c_size_t]
    L.lb2_free.argtypes = [vp]
    L.lb2_malloc_host.restype = vp
    L.lb2_malloc_host.argtypes = [C.c_size_t]
    L.lb2_free_host.argtypes = [vp]
    L.lb2_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.lb2_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.lb2_memset.argtypes = [vp, vp, ci, C.c_size_t]
    L.lb2_op_stencil.restype = vp
    L.lb2_op_stencil.argtypes = [C.c_char, i64, i64, i64, dbl, dbl, vp]
    L.lb2_op_csr.restype = vp
    L.lb2_op_csr.argtypes = [C.c_char, i64, vp, vp, vp]
    L.lb2_op_diag.restype = vp
    L.lb2_op_diag.argtypes = [C.c_char, i64, vp]
    L.lb2_op_bdg.restype = vp
    L.lb2_op_bdg.argtypes = [C.c_char, i64, i64, i64, dbl, dbl, dbl, dbl, dbl]
    L.lb2_op_csr32.restype = vp
    L.lb2_op_csr32.argtypes = [C.c_char, i64, vp, vp, vp]
    L.lb2_op_csr_from_mtx.restype = vp
    L.lb2_op_csr_from_mtx.argtypes = [C.c_char, C.c_char_p]
    L.lb2_op_dense.restype = vp
    L.lb2_op_dense.argtypes = [C.c_char, i64, vp]
    L.lb2_write_mtx.argtypes = [C.c_char_p, C.c_char, i64, i64, vp, i64]
    L.lb2_op_set_halo.argtypes = [vp, vp, vp, i64]
    L.lb2_op_bdg_slab.restype = vp
    L.lb2_op_bdg_slab.argtypes = [C.c_char, i64, i64, i64, i64, i64, dbl, dbl, dbl, dbl, dbl]
    L.lb2_op_spec_hi.restype = dbl
    L.lb2_op_spec_hi.argtypes = [vp]
    L.lb2_op_csr_slab.restype = vp
    L.lb2_op_csr_slab.argtypes = [C.c_char, i64, i64, i64, vp, vp, vp]
    L.lb2_op_device.restype = vp
    L.lb2_op_device.argtypes = [C.c_char, i64, vp, vp, dbl]
    L.lb2_op_chebyshev.restype = vp
    L.lb2_op_chebyshev.argtypes = [C.c_char, vp, ci, dbl, dbl]
    L.lb2_op_chebyshev_mixed.restype = vp
    L.lb2_op_chebyshev_mixed.argtypes = [C.c_char, vp, ci, dbl, dbl]
    L.lb2_op_destroy.argtypes = [vp]
    L.lb2_op_apply.argtypes = [vp, vp, C.c_char, ci, vp, i64, vp, i64]
    L.lb2_solver_create.restype = vp
    L.lb2_solver_create.argtypes = [vp, C.c_char, vp, ci]
    for f in ("init", "finish"):
        getattr(L, f"lb2_solver_{f}").argtypes = [vp]
    L.lb2_solver_step.argtypes = [vp, ci]
    L.lb2_solver_prepare.argtypes = [vp]
    L.lb2_solver_arena.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.lb2_solver_set_peers.argtypes = [vp, vp, vp]
    L.lb2_op_stencil_slab.restype = vp
    L.lb2_op_stencil_slab.argtypes = [C.c_char, i64, i64, i64, i64, i64, dbl, dbl, vp]
    L.lb2_solver_destroy.argtypes = [vp]
    L.lb2_solver_set_device_x0.argtypes = [vp, u64]
    L.lb2_solver_set_device_io.argtypes = [vp, vp, vp]
    L.lb2_solver_stat_name.restype = C.c_char_p
    L.lb2_solver_stat_name.argtypes = [ci]
    L.lb2_solver_stat.restype = dbl
    L.lb2_solver_stat.argtypes = [vp, ci]
    L.lb2_solver_stat_work.restype = dbl
    L.lb2_solver_stat_work.argtypes = [vp, ci]
    L.lb2_solver_stat_calls.restype = C.c_ulonglong
    L.lb2_solver_stat_calls.argtypes = [vp, ci]
    L.lb2_solver_reset_stats.argtypes = [vp]
    L.lb2_solver_set_option.argtypes = [vp, C.c_char_p, ci]
    L.lb2_solver_info.restype = dbl
    L.lb2_solver_info.argtypes = [vp, C.c_char_p]
    L.lb2_set_num_gpus.argtypes = [ci]
    L.lb2_last_num_gpus.restype = ci
    L.lb2_last_status.restype = ci
    L.lb2_solver_results.argtypes = [vp, C.POINTER(dbl), ci, C.POINTER(dbl), ci]
    L.lb2_solver_state.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(ci)]
    L.lb2_comm_unique_id.argtypes = [vp, C.c_char_p]
    L.lb2_ctx_attach_comm.argtypes = [vp, ci, ci, vp, C.c_char_p]
    L.lb2_ctx_detach_comm.argtypes = [vp]
    L.lb2_comm_allreduce.argtypes = [vp, vp, C.c_size_t, ci]
    L.lb2_ipc_get_handle.argtypes = [vp, vp]
    L.lb2_ipc_open_handle.restype = vp
    L.lb2_ipc_open_handle.argtypes = [vp]
    L.lb2_ipc_close_handle.argtypes = [vp]
    for p in "sdcz":
        getattr(L, f"lb2_{p}_state_alloc").restype = vp
        getattr(L, f"lb2_{p}_state_alloc").argtypes = [u64, u64, u64, ci]
        getattr(L, f"lb2_{p}_state_free").argtypes = [vp]
        getattr(L, f"{p}_lobpcg").argtypes = [vp]
        getattr(L, f"{p}_ilobpcg").argtypes = [vp]
        getattr(L, f"lb2_{p}_gram").argtypes = [vp, i64, ci, ci, vp, i64, vp, i64, vp, ci, ci]
        getattr(L, f"lb2_{p}_gram_cols").argtypes = [vp, i64, ci, ci, vp, i64, vp, i64, vp, ci, vp, i64, vp, ci, ci]
        getattr(L, f"lb2_{p}_tall_nn").argtypes = [vp, i64, ci, ci, vp, vp, i64, vp, ci, vp, vp, i64]
        getattr(L, f"lb2_{p}_residual").argtypes = [vp, i64, ci, vp, i64, vp, i64, vp, vp, i64, vp]
        getattr(L, f"lb2_{p}_col_sumsq").argtypes = [vp, i64, ci, vp, i64, vp]
        getattr(L, f"lb2_{p}_fill_uniform").argtypes = [vp, i64, ci, vp, i64, u64, i64, i64]
        getattr(L, f"lb2_{p}_spmm_stencil").argtypes = [vp, i64, i64, i64, dbl, dbl, vp, ci, vp, i64, vp, i64]
        getattr(L, f"lb2_{p}_spmm_stencil_halo").argtypes = [vp, i64, i64, i64, dbl, dbl, vp, vp, vp, i64, ci, vp, i64, vp, i64]
        getattr(L, f"lb2_{p}_spmm_csr").argtypes = [vp, i64, vp, vp, vp, ci, vp, i64, vp, i64]
        getattr(L, f"lb2_{p}_spmm_diag").argtypes = [vp, i64, vp, ci, vp, i64, vp, i64]


def _ck(rc, what):
    if rc != 0:
        raise LobpcgB200Error(f"{what} failed with code {rc}")


# --------------------------------------------------------------------------------------------------- context
class Context:
    def __init__(self, device: int = -1, stream: int | None = None):
        self.h = lib().lb2_ctx_create(device, stream)
        if not self.h:
            raise LobpcgB200Error("lb2_ctx_create failed (no CUDA device?)")

    def sync(self):
        _ck(lib().lb2_ctx_sync(self.h), "lb2_ctx_sync")

    def set_option(self, key: str, value: int):
        _ck(lib().lb2_ctx_set_option(self.h, key.encode(), int(value)), f"set_option({key})")

    def oz_stats(self):
        """{split_ms, mma_ms, reduce_ms, calls} of the int8 tensor-path Gram since the last query."""
        out = (C.c_double * 4)()
        _ck(lib().lb2_ctx_oz_stats(self.h, out), "lb2_ctx_oz_stats")
        return dict(split_ms=out[0], mma_ms=out[1], reduce_ms=out[2], calls=int(out[3]))

    def trim(self):
        """Free the solver arena the context keeps for reuse by its next solve."""
        _ck(lib().lb2_ctx_trim(self.h), "lb2_ctx_trim")

    @property
    def launches(self) -> int:
        return int(lib().lb2_ctx_launches(self.h))

    def close(self):
        if self.h:
            lib().lb2_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceArray:
    """Column-major device block (n x nc, leading dimension ld >= n) owned through lb2_malloc."""

    def __init__(self, shape, dtype, ld=None):
        self.shape = tuple(int(v) for v in shape)
        self.dtype = np.dtype(dtype)
        n = self.shape[0]
        nc = self.shape[1] if len(self.shape) > 1 else 1
        self.ld = int(ld) if ld else max(n, 1)
        self.nbytes = self.ld * max(nc, 1) * self.dtype.itemsize
        self.ptr = lib().lb2_malloc(self.nbytes)
        if not self.ptr:
            raise LobpcgB200Error(f"device allocation of {self.nbytes} bytes failed")

    @classmethod
    def from_numpy(cls, ctx: Context, a: np.ndarray, ld=None):
        a = np.asarray(a)
        d = cls(a.shape, a.dtype, ld)
        d.upload(ctx, a)
        return d

    def upload(self, ctx: Context, a: np.ndarray):
        a = np.asarray(a, dtype=self.dtype)
        n = self.shape[0]
        nc = self.shape[1] if len(self.shape) > 1 else 1
        buf = np.zeros((self.ld, nc), dtype=self.dtype, order="F")
        buf[:n, :] = a.reshape((n, nc), order="F") if a.ndim == 1 else a
        _ck(lib().lb2_memcpy_h2d(ctx.h, self.ptr, buf.ctypes.data, buf.nbytes), "h2d")

    def numpy(self, ctx: Context) -> np.ndarray:
        n = self.shape[0]
        nc = self.shape[1] if len(self.shape) > 1 else 1
        buf = np.empty((self.ld, nc), dtype=self.dtype, order="F")
        _ck(lib().lb2_memcpy_d2h(ctx.h, buf.ctypes.data, self.ptr, buf.nbytes), "d2h")
        out = np.asfortranarray(buf[:n, :])
        return out[:, 0].copy() if len(self.shape) == 1 else out

    def rows(self, row0: int, nrows: int) -> "DeviceArray":
        """Non-owning view of rows [row0, row0+nrows) of every column (same leading dimension)."""
        v = object.__new__(DeviceArray)
        v.shape = (int(nrows),) + tuple(self.shape[1:])
        v.dtype, v.ld = self.dtype, self.ld
        v.ptr = self.ptr + int(row0) * self.dtype.itemsize
        v.nbytes = 0
        v._owner = self
        return v

    def cols(self, col0: int, ncols: int) -> "DeviceArray":
        """Non-owning view of columns [col0, col0+ncols) (same leading dimension)."""
        v = object.__new__(DeviceArray)
        v.shape = (self.shape[0], int(ncols))
        v.dtype, v.ld = self.dtype, self.ld
        v.ptr = self.ptr + int(col0) * self.ld * self.dtype.itemsize
        v.nbytes = 0
        v._owner = self
        return v

    def zero(self, ctx: Context):
        _ck(lib().lb2_memset(ctx.h, self.ptr, 0, self.nbytes), "memset")

    def free(self):
        if self.ptr and not hasattr(self, "_owner"):
            lib().lb2_free(self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _scalar(dtype, v):
    return np.array([v], dtype=dtype)


# --------------------------------------------------------------------------------------------------- kernels
def gram(ctx, A: DeviceArray, B: DeviceArray, upper=False) -> DeviceArray:
    p = PREFIX[A.dtype]
    n, ma = A.shape
    mb = B.shape[1]
    G = DeviceArray((ma, mb), A.dtype)
    _ck(getattr(lib(), f"lb2_{p}_gram")(ctx.h, n, ma, mb, A.ptr, A.ld, B.ptr, B.ld, G.ptr, G.ld, int(upper)), "gram")
    return G


def gram_cols(ctx, S: DeviceArray, W0: DeviceArray, W1: "DeviceArray | None" = None, tri_c0: int = -1):
    """(S^H W0, S^H W1) through lb2_<p>_gram_cols; entries strictly below the diagonal of the Hermitian block that starts
    at row tri_c0 may be left as they were (the outputs are zero-initialised here)."""
    p = PREFIX[S.dtype]
    n, m = S.shape
    nw = W0.shape[1]
    G0 = DeviceArray((m, nw), S.dtype); G0.zero(ctx)
    G1 = None
    if W1 is not None:
        G1 = DeviceArray((m, nw), S.dtype); G1.zero(ctx)
    _ck(getattr(lib(), f"lb2_{p}_gram_cols")(ctx.h, n, m, nw, S.ptr, S.ld, W0.ptr, W0.ld, G0.ptr, G0.ld,
                                             W1.ptr if W1 is not None else None, W1.ld if W1 is not None else 0,
                                             G1.ptr if G1 is not None else None, G1.ld if G1 is not None else 0,
                                             int(tri_c0)), "gram_cols")
    return G0, G1


def tall_nn(ctx, S: DeviceArray, Cm: DeviceArray, Out: DeviceArray, alpha=1.0, beta=0.0):
    p = PREFIX[S.dtype]
    n, kd = S.shape
    nb = Cm.shape[1]
    a, b = _scalar(S.dtype, alpha), _scalar(S.dtype, beta)
    _ck(getattr(lib(), f"lb2_{p}_tall_nn")(ctx.h, n, kd, nb, a.ctypes.data, S.ptr, S.ld, Cm.ptr, Cm.ld,
                                           b.ctypes.data, Out.ptr, Out.ld), "tall_nn")
    return Out


def residual(ctx, AX: DeviceArray, BX: DeviceArray, lam: DeviceArray, write=True, norms=True):
    p = PREFIX[AX.dtype]
    n, nc = AX.shape
    W = DeviceArray((n, nc), AX.dtype) if write else None
    ss = DeviceArray((nc,), REAL[p]) if norms else None
    _ck(getattr(lib(), f"lb2_{p}_residual")(ctx.h, n, nc, AX.ptr, AX.ld, BX.ptr, BX.ld, lam.ptr,
                                            W.ptr if W else None, W.ld if W else 0, ss.ptr if ss else None), "residual")
    return W, ss


def col_sumsq(ctx, X: DeviceArray) -> DeviceArray:
    p = PREFIX[X.dtype]
    n, nc = X.shape
    ss = DeviceArray((nc,), REAL[p])
    _ck(getattr(lib(), f"lb2_{p}_col_sumsq")(ctx.h, n, nc, X.ptr, X.ld, ss.ptr), "col_sumsq")
    return ss


def fill_uniform(ctx, n, nc, dtype, seed, n_global=None, row0=0, ld=None) -> DeviceArray:
    p = PREFIX[np.dtype(dtype)]
    X = DeviceArray((n, nc), dtype, ld)
    _ck(getattr(lib(), f"lb2_{p}_fill_uniform")(ctx.h, n, nc, X.ptr, X.ld, seed, n_global or n, row0), "fill_uniform")
    return X


# --------------------------------------------------------------------------------------------------- operators
class LinOpStruct(C.Structure):
    """LinearOperator_<p>_t (reference include/lobpcg/linop.h:20-26)."""
    _fields_ = [("rows", C.c_uint64), ("cols", C.c_uint64), ("matvec", C.c_void_p), ("cleanup", C.c_void_p),
                ("ctx", C.c_void_p)]


class LinOpCtx(C.Structure):
    _fields_ = [("data", C.c_void_p), ("data_size", C.c_size_t)]


class LinOp:
    """Handle to a LinearOperator_<p>_t*.  Built-in operators live on the device."""

    def __init__(self, handle, prefix, n, builtin=True, keep=()):
        if not handle:
            raise LobpcgB200Error("operator construction failed")
        self.handle, self.prefix, self.n, self.builtin, self._keep = handle, prefix, n, builtin, keep

    def apply(self, ctx: Context, X: DeviceArray, Y: DeviceArray | None = None) -> DeviceArray:
        nc = X.shape[1] if len(X.shape) > 1 else 1
        if Y is None:
            Y = DeviceArray(X.shape, X.dtype, X.ld)
        _ck(lib().lb2_op_apply(ctx.h, self.handle, self.prefix.encode(), nc, X.ptr, X.ld, Y.ptr, Y.ld), "op_apply")
        return Y

    def close(self):
        if self.handle and self.builtin:
            lib().lb2_op_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def stencil_op(grid, dtype, cdiag=None, coff=-1.0, potential=None) -> LinOp:
    p = PREFIX[np.dtype(dtype)]
    g = tuple(int(v) for v in grid) + (1,) * (3 - len(grid))
    cdiag = 2.0 * len(grid) if cdiag is None else cdiag
    pot = None if potential is None else np.ascontiguousarray(potential, dtype=REAL[p])
    h = lib().lb2_op_stencil(p.encode(), g[0], g[1], g[2], cdiag, coff, pot.ctypes.data if pot is not None else None)
    return LinOp(h, p, g[0] * g[1] * g[2])


def stencil_slab_op(grid, z0, gz_local, dtype, cdiag=None, coff=-1.0, potential_local=None) -> LinOp:
    """Rank-local z-slab [z0, z0+gz_local) of a 3-D stencil (row-partitioned multi-GPU runs)."""
    p = PREFIX[np.dtype(dtype)]
    gx, gy, gz = (int(v) for v in grid)
    cdiag = 6.0 if cdiag is None else cdiag
    pot = None if potential_local is None else np.ascontiguousarray(potential_local, dtype=REAL[p])
    h = lib().lb2_op_stencil_slab(p.encode(), gx, gy, gz_local, gz, z0, cdiag, coff,
                                  pot.ctypes.data if pot is not None else None)
    return LinOp(h, p, gx * gy * gz_local)


def bdg_slab_op(grid, z0, gz_local, dtype, shift, d, cdiag=6.0, coff=-1.0) -> LinOp:
    """Rank-local part of the BdG operator (lb2_op_bdg_slab): local rows = [u z-slab ; v z-slab]."""
    p = PREFIX[np.dtype(dtype)]
    gx, gy, gz = (int(v) for v in grid)
    h = lib().lb2_op_bdg_slab(p.encode(), gx, gy, int(gz_local), gz, int(z0), cdiag, coff, shift,
                              float(np.real(d)), float(np.imag(d)))
    return LinOp(h, p, 2 * gx * gy * int(gz_local))


def csr_slab_op(n_global: int, row0: int, rowptr_local, col_global, val) -> LinOp:
    """Row block [row0, row0 + n_local) of a CSR matrix for a row-partitioned run (lb2_op_csr_slab): ``rowptr_local``
    starts at 0, ``col_global`` holds global column indices."""
    val = np.ascontiguousarray(val)
    p = PREFIX[val.dtype]
    rp = np.ascontiguousarray(rowptr_local, dtype=np.int64)
    col = np.ascontiguousarray(col_global, dtype=np.int32)
    n_local = len(rp) - 1
    h = lib().lb2_op_csr_slab(p.encode(), int(n_global), int(row0), n_local, rp.ctypes.data, col.ctypes.data,
                              val.ctypes.data)
    return LinOp(h, p, n_local)


def set_halo(op: LinOp, lo, hi, ld: int):
    """Neighbour data (device pointers or None) for a stand-alone apply of a row-block operator (lb2_op_set_halo)."""
    _ck(lib().lb2_op_set_halo(op.handle, lo, hi, int(ld)), "lb2_op_set_halo")


def spec_hi(op: LinOp) -> float:
    """Upper bound of the operator's spectrum recorded at construction (0 = unknown)."""
    return float(lib().lb2_op_spec_hi(op.handle))


def stencil_halo_apply(ctx, grid_local, X: DeviceArray, halo_lo, halo_hi, halo_ld, cdiag=6.0, coff=-1.0) -> DeviceArray:
    """Kernel-level stencil on one z-slab with explicit halo-plane device pointers (ints or None)."""
    p = PREFIX[X.dtype]
    gx, gy, gz = grid_local
    Y = DeviceArray(X.shape, X.dtype)
    _ck(getattr(lib(), f"lb2_{p}_spmm_stencil_halo")(ctx.h, gx, gy, gz, cdiag, coff, None, halo_lo, halo_hi, halo_ld,
                                                     X.shape[1], X.ptr, X.ld, Y.ptr, Y.ld), "spmm_stencil_halo")
    return Y


def csr_op(rowptr, col, val) -> LinOp:
    val = np.ascontiguousarray(val)
    p = PREFIX[val.dtype]
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    n = len(rowptr) - 1
    h = lib().lb2_op_csr(p.encode(), n, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data)
    return LinOp(h, p, n)


def diag_op(d, dtype) -> LinOp:
    p = PREFIX[np.dtype(dtype)]
    d = np.ascontiguousarray(d, dtype=REAL[p])
    return LinOp(lib().lb2_op_diag(p.encode(), len(d), d.ctypes.data), p, len(d))


def bdg_op(grid, dtype, shift, d, cdiag=None, coff=-1.0) -> LinOp:
    p = PREFIX[np.dtype(dtype)]
    g = tuple(int(v) for v in grid) + (1,) * (3 - len(grid))
    cdiag = 2.0 * len(grid) if cdiag is None else cdiag
    d = complex(d)
    return LinOp(lib().lb2_op_bdg(p.encode(), g[0], g[1], g[2], cdiag, coff, shift, d.real, d.imag), p,
                 2 * g[0] * g[1] * g[2])


def mtx_op(path, dtype) -> LinOp:
    """CSR operator read from a Matrix Market coordinate file (lb2_op_csr_from_mtx)."""
    p = PREFIX[np.dtype(dtype)]
    h = lib().lb2_op_csr_from_mtx(p.encode(), str(path).encode())
    if not h:
        raise LobpcgB200Error(f"cannot read {path}")
    op = LinOpStruct.from_address(h)
    return LinOp(h, p, int(op.rows))


def dense_op(A) -> LinOp:
    """Device-resident dense operator (lb2_op_dense): A is an n x n numpy matrix."""
    A = np.asfortranarray(A)
    p = PREFIX[A.dtype]
    return LinOp(lib().lb2_op_dense(p.encode(), A.shape[0], A.ctypes.data), p, A.shape[0])


def write_mtx(path, a: np.ndarray):
    """Write a host vector / block as a Matrix Market dense array file (lb2_write_mtx)."""
    a = np.asfortranarray(a if np.ndim(a) == 2 else np.asarray(a).reshape(-1, 1))
    _ck(lib().lb2_write_mtx(str(path).encode(), PREFIX[a.dtype].encode(), a.shape[0], a.shape[1], a.ctypes.data,
                            max(a.shape[0], 1)), "lb2_write_mtx")


MATMAT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p)


def device_op(n: int, dtype, fn, spec_hi: float = 0.0) -> LinOp:
    """Caller-supplied block operator on device pointers (lb2_op_device): ``fn(ncols, X_ptr, ldx, Y_ptr, ldy, stream)``
    enqueues Y = Op X on ``stream`` and returns 0."""
    p = PREFIX[np.dtype(dtype)]
    cb = MATMAT_FN(lambda user, nc, X, ldx, Y, ldy, stream: int(fn(nc, X, ldx, Y, ldy, stream)))
    h = lib().lb2_op_device(p.encode(), n, C.cast(cb, C.c_void_p), None, spec_hi)
    return LinOp(h, p, n, keep=(cb, fn))


def chebyshev_op(A: LinOp, degree: int, lo: float = 0.0, hi: float = 0.0, mixed: bool = False) -> LinOp:
    """Built-in preconditioner T = p(A) (lb2_op_chebyshev): `degree` Chebyshev steps for A y = x on [lo, hi];
    mixed=True evaluates it in float / complex float inside a double solve (lb2_op_chebyshev_mixed)."""
    fn = lib().lb2_op_chebyshev_mixed if mixed else lib().lb2_op_chebyshev
    h = fn(A.prefix.encode(), A.handle, int(degree), float(lo), float(hi))
    return LinOp(h, A.prefix, A.n, keep=(A,))


def host_op(n, dtype, fn) -> LinOp:
    """Foreign operator: a host callback y = fn(x) on single vectors, exactly the reference's
    matvec_func_<p>_t (linop.h:15-17).  The solver stages columns through host memory for these."""
    dtype = np.dtype(dtype)
    p = PREFIX[dtype]
    MV = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)

    def _mv(op, x, y):
        xv = np.ctypeslib.as_array(C.cast(x, C.POINTER(C.c_byte)), shape=(n * dtype.itemsize,)).view(dtype)
        yv = np.ctypeslib.as_array(C.cast(y, C.POINTER(C.c_byte)), shape=(n * dtype.itemsize,)).view(dtype)
        yv[:] = fn(xv)

    cb = MV(_mv)
    ctx = LinOpCtx(None, 0)
    st = LinOpStruct(n, n, C.cast(cb, C.c_void_p), None, C.cast(C.pointer(ctx), C.c_void_p))
    return LinOp(C.addressof(st), p, n, builtin=False, keep=(cb, ctx, st))


# --------------------------------------------------------------------------------------------------- solver
def _state_struct(prefix):
    rt = C.c_float if prefix in "sc" else C.c_double
    vp = C.c_void_p

    class State(C.Structure):
        _fields_ = [("S", vp), ("Cx", vp), ("Cp", vp), ("AX", vp), ("AS", vp), ("BS", vp), ("eigVals", vp),
                    ("resNorm", vp), ("signature", vp), ("wrk1", vp), ("wrk2", vp), ("wrk3", vp), ("wrk4", vp),
                    ("rr_D", vp), ("rr_eigvals", vp), ("rr_tau", vp), ("rr_VR", vp), ("rr_sig", vp),
                    ("rr_indices", vp), ("rr_ggev", vp), ("implicit_product_update", C.c_int8),
                    ("verbosity", C.c_int8), ("iter", C.c_uint64), ("nev", C.c_uint64), ("converged", C.c_uint64),
                    ("size", C.c_uint64), ("sizeSub", C.c_uint64), ("maxIter", C.c_uint64), ("tol", rt),
                    ("A", vp), ("B", vp), ("T", vp)]

    return State


class SolverState:
    """Owns a ``<p>_lobpcg_t`` allocated by lb2_<p>_state_alloc (host buffers)."""

    def __init__(self, dtype, n, nev, k, indefinite=False):
        self.dtype = np.dtype(dtype)
        self.prefix = PREFIX[self.dtype]
        self.n, self.nev, self.k, self.indefinite = int(n), int(nev), int(k), bool(indefinite)
        self.ptr = getattr(lib(), f"lb2_{self.prefix}_state_alloc")(self.n, self.nev, self.k, int(indefinite))
        if not self.ptr:
            raise LobpcgB200Error("state allocation failed")
        self.st = _state_struct(self.prefix).from_address(self.ptr)

    def X(self) -> np.ndarray:
        """View of alg->S[0 : n*k) as an (n,k) Fortran array."""
        buf = (C.c_byte * (self.n * self.k * self.dtype.itemsize)).from_address(self.st.S)
        return np.frombuffer(buf, dtype=self.dtype).reshape((self.n, self.k), order="F")

    def eigvals(self):
        buf = (C.c_byte * (self.k * REAL[self.prefix].itemsize)).from_address(self.st.eigVals)
        return np.frombuffer(buf, dtype=REAL[self.prefix]).copy()

    def resnorm(self):
        buf = (C.c_byte * (self.k * REAL[self.prefix].itemsize)).from_address(self.st.resNorm)
        return np.frombuffer(buf, dtype=REAL[self.prefix]).copy()

    def signature(self):
        if not self.st.signature:
            return None
        buf = (C.c_byte * (3 * self.k)).from_address(self.st.signature)
        return np.frombuffer(buf, dtype=np.int8).copy()

    def free(self):
        if self.ptr:
            getattr(lib(), f"lb2_{self.prefix}_state_free")(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _setup(A, X0, n, k, nev, dtype, tol, max_iter, B, T, indefinite, verbosity):
    st = SolverState(dtype, n, nev, k, indefinite)
    st.st.A = A.handle
    st.st.B = B.handle if B is not None else None
    st.st.T = T.handle if T is not None else None
    st.st.maxIter = int(max_iter)
    st.st.tol = float(tol)
    st.st.verbosity = int(verbosity)
    if X0 is not None:
        st.X()[:, :] = X0
    return st


def lobpcg(A: LinOp, X0: np.ndarray, nev: int, tol: float, max_iter: int, B: LinOp | None = None,
           T: LinOp | None = None, indefinite: bool = False, verbosity: int = 0):
    """Drop-in call path: fill a ``<p>_lobpcg_t`` with HOST buffers and call ``<p>_lobpcg`` / ``<p>_ilobpcg``."""
    X0 = np.asarray(X0)
    n, k = X0.shape
    st = _setup(A, X0, n, k, nev, X0.dtype, tol, max_iter, B, T, indefinite, verbosity)
    fn = getattr(lib(), f"{st.prefix}_{'i' if indefinite else ''}lobpcg")
    fn(st.ptr)
    out = dict(eig=st.eigvals(), res=st.resnorm(), X=np.array(st.X(), order="F", copy=True),
               iter=int(st.st.iter), converged=int(st.st.converged), sig=st.signature(), status=int(lib().lb2_last_status()))
    st.free()
    return out


class Solver:
    """Resumable device-resident solver (lb2_solver_*): init / step(k) / finish on one state struct."""

    def __init__(self, ctx: Context, A: LinOp, n: int, k: int, nev: int, dtype, tol: float, max_iter: int,
                 B: LinOp | None = None, T: LinOp | None = None, X0: np.ndarray | None = None,
                 device_seed: int | None = None, indefinite: bool = False, verbosity: int = 0):
        self.ctx = ctx
        self.ops = (A, B, T)
        self.state_ = _setup(A, X0, n, k, nev, dtype, tol, max_iter, B, T, indefinite, verbosity)
        self.h = lib().lb2_solver_create(ctx.h, self.state_.prefix.encode(), self.state_.ptr, int(indefinite))
        if not self.h:
            raise LobpcgB200Error("lb2_solver_create failed")
        if device_seed is not None:
            lib().lb2_solver_set_device_x0(self.h, int(device_seed))

    def set_device_io(self, x0: "DeviceArray | None" = None, x_out: "DeviceArray | None" = None):
        """Device-pointer fast path: read X0 from / write the eigenvectors to device blocks (no host copies)."""
        self._dev_io = (x0, x_out)
        _ck(lib().lb2_solver_set_device_io(self.h, x0.ptr if x0 is not None else None,
                                           x_out.ptr if x_out is not None else None), "lb2_solver_set_device_io")

    def set_option(self, key: str, value: int):
        _ck(lib().lb2_solver_set_option(self.h, key.encode(), int(value)), f"lb2_solver_set_option({key})")

    def info(self, key: str) -> float:
        return float(lib().lb2_solver_info(self.h, key.encode()))

    def prepare(self):
        _ck(lib().lb2_solver_prepare(self.h), "lb2_solver_prepare")

    def arena(self):
        ptr, nbytes = C.c_void_p(0), C.c_size_t(0)
        _ck(lib().lb2_solver_arena(self.h, C.byref(ptr), C.byref(nbytes)), "lb2_solver_arena")
        return ptr.value, nbytes.value

    def set_peers(self, lo, hi):
        _ck(lib().lb2_solver_set_peers(self.h, lo, hi), "lb2_solver_set_peers")

    def init(self):
        _ck(lib().lb2_solver_init(self.h), "lb2_solver_init")

    def step(self, max_steps: int) -> int:
        rc = lib().lb2_solver_step(self.h, int(max_steps))
        if rc < 0:
            raise LobpcgB200Error(f"lb2_solver_step failed with code {rc}")
        return rc

    def finish(self):
        _ck(lib().lb2_solver_finish(self.h), "lb2_solver_finish")
        s = self.state_
        # X is copied: the state's host block is freed by close() / garbage collection
        return dict(eig=s.eigvals(), res=s.resnorm(), X=np.array(s.X(), order="F", copy=True), iter=int(s.st.iter),
                    converged=int(s.st.converged), sig=s.signature())

    def progress(self):
        it, cv, uo = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        lib().lb2_solver_state(self.h, C.byref(it), C.byref(cv), C.byref(uo))
        return dict(iter=it.value, converged=cv.value, use_ortho=uo.value)

    def results(self):
        """(eig[0:k], res[0:nev]) of the last pass without downloading the eigenvectors."""
        k, nev = self.state_.k, self.state_.nev
        e, r = (C.c_double * k)(), (C.c_double * nev)()
        _ck(lib().lb2_solver_results(self.h, e, k, r, nev), "lb2_solver_results")
        return np.array(e), np.array(r)

    def stats(self) -> dict:
        """{phase: dict(ms, work, calls)}; work = algorithmic flops (gram, tall_nn) or bytes (spmm, residual)."""
        L = lib()
        out = {}
        for i in range(L.lb2_solver_num_stats()):
            name = L.lb2_solver_stat_name(i).decode().replace("_ms", "")
            out[name] = dict(ms=L.lb2_solver_stat(self.h, i), work=L.lb2_solver_stat_work(self.h, i),
                             calls=int(L.lb2_solver_stat_calls(self.h, i)))
        return out

    def reset_stats(self):
        lib().lb2_solver_reset_stats(self.h)

    def close(self):
        if self.h:
            lib().lb2_solver_destroy(self.h)
            self.h = None
        if self.state_:
            self.state_.free()
            self.state_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
