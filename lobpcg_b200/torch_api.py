"""torch-facing wrapper of the device-pointer fast path ("caller-side bindings").

CUDA tensors go in and come out; nothing is staged through the host.  torch is plumbing here: it owns the memory and
the stream, the work is done by the library's kernels through ``lb2_solver_set_device_io`` (include/lobpcg_b200.h) on a
context bound to torch's current stream, so solver launches are ordered with the caller's torch work like any other op.

A block vector is an ``(n, k)`` tensor with strides ``(1, n)`` — the reference's column-major layout
(``include/lobpcg/blas_wrapper.h:5-8``); ``torch.empty(k, n).T`` has it, other layouts are copied once.
"""
from __future__ import annotations

import numpy as np
import torch

from . import api

_TYPESTR = {np.dtype(np.float32): "<f4", np.dtype(np.float64): "<f8", np.dtype(np.complex64): "<c8",
            np.dtype(np.complex128): "<c16"}
_NP = {torch.float32: np.float32, torch.float64: np.float64, torch.complex64: np.complex64,
       torch.complex128: np.complex128}
_contexts: dict = {}


def _context(device: torch.device) -> api.Context:
    """One library context per (device, torch stream)."""
    # torch reports its legacy default stream as 0; the library takes 0 as "create your own stream", so the default
    # stream is named by its explicit handle cudaStreamLegacy (0x1)
    stream = torch.cuda.current_stream(device).cuda_stream or 1
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    ctx = _contexts.get(key)
    if ctx is None:
        ctx = _contexts[key] = api.Context(key[0], stream)
    return ctx


def _column_major(X: torch.Tensor) -> torch.Tensor:
    n, k = X.shape
    if X.stride() == (1, n):
        return X
    return X.T.contiguous().T


class _View:
    """Borrowed device block exposed through ``__cuda_array_interface__`` (shape (nc, n), row stride ld)."""

    def __init__(self, ptr, nc, n, ld, dtype, read_only):
        item = np.dtype(dtype).itemsize
        self.__cuda_array_interface__ = {"shape": (int(nc), int(n)), "strides": (int(ld) * item, item),
                                         "typestr": _TYPESTR[np.dtype(dtype)], "data": (int(ptr), bool(read_only)),
                                         "version": 2}


class _Block:
    """What ``Solver.set_device_io`` needs from a device block: the pointer."""

    def __init__(self, t: torch.Tensor):
        self.t, self.ptr = t, t.data_ptr()


def torch_op(n: int, dtype: torch.dtype, fn, spec_hi: float = 0.0) -> api.LinOp:
    """Operator from a torch function on device blocks: ``fn(X, Y)`` writes ``Y = Op X`` where X and Y are ``(n, nc)``
    tensors viewing the solver's own memory (strides (1, ld)).  Uses ``lb2_op_device``; run the solve through
    :func:`lobpcg` of this module so that the callback's torch ops and the library's kernels share a stream."""
    npdt = _NP[dtype]

    def matmat(nc, X, ldx, Y, ldy, stream):
        with torch.no_grad():
            Xt = torch.as_tensor(_View(X, nc, n, ldx, npdt, False), device="cuda").T
            Yt = torch.as_tensor(_View(Y, nc, n, ldy, npdt, False), device="cuda").T
            fn(Xt, Yt)
        return 0

    return api.device_op(n, npdt, matmat, spec_hi)


def lobpcg(A: api.LinOp, X0: torch.Tensor, nev: int, tol: float = 1e-8, max_iter: int = 1000,
           B: api.LinOp | None = None, T: api.LinOp | None = None, indefinite: bool = False):
    """Solve for the ``nev`` smallest eigenpairs of (A, B) from the CUDA block ``X0`` (n x sizeSub, sizeSub >= 2 nev).
    Returns ``(eigvals, X, info)``: eigenvalues as a tensor on X0's device, eigenvectors as an (n, sizeSub) column-major
    tensor, ``info`` = dict(iter, converged, res)."""
    if not X0.is_cuda:
        raise ValueError("X0 must be a CUDA tensor (host buffers go through lobpcg_b200.api.lobpcg)")
    n, k = X0.shape
    npdt = _NP[X0.dtype]
    with torch.cuda.device(X0.device):
        ctx = _context(X0.device)
        Xin = _column_major(X0)
        Xout = torch.empty((k, n), dtype=X0.dtype, device=X0.device).T
        s = api.Solver(ctx, A, n, k, nev, npdt, tol, max_iter, B=B, T=T, indefinite=indefinite)
        try:
            s.set_device_io(_Block(Xin), _Block(Xout))
            s.init()
            s.step(max_iter + 1)
            r = s.finish()
        finally:
            s.close()
    eig = torch.as_tensor(np.array(r["eig"]), device=X0.device)
    return eig, Xout, dict(iter=r["iter"], converged=r["converged"], res=np.array(r["res"]))
