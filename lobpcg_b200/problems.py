"""Synthetic problem generators shared by tests, bench and the oracle harness.

Everything here is host-side numpy and deterministic (portable across machines):

* ``splitmix_uniform``  counter-based splitmix64 -> uniform [-0.5, 0.5); the CUDA library implements the
  same generator (``lb2_fill_uniform_*`` in csrc/elementwise.cu) so X0 can be produced on the device
  bit-identically.  Replaces the reference's libc ``rand()`` fill
  (src/residual/estimate_norm_impl.inc:19-35), which is not portable.
* Dirichlet stencil Laplacians: 1-D/2-D/3-D, diagonal 2*dim, off-diagonals -1, natural
  ordering with x fastest; CSR with int32 ascending column indices.
* analytic spectra for those Laplacians and the BdG-style pencil of config C4.
"""
from __future__ import annotations

import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed: int, start: int, count: int) -> np.ndarray:
    """z_i = mix(seed + (start+i+1)*GOLDEN) for i in [0,count) as uint64."""
    with np.errstate(over="ignore"):
        idx = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def splitmix_uniform(seed: int, count: int, dtype=np.float64, start: int = 0) -> np.ndarray:
    """uniform [-0.5,0.5): top 53 bits (f64) or top 24 bits (f32) of splitmix64.

    Complex dtypes consume two consecutive counters per element (re, im)."""
    dtype = np.dtype(dtype)
    if dtype.kind == "c":
        rt = np.float32 if dtype == np.complex64 else np.float64
        r = splitmix_uniform(seed, 2 * count, rt, 2 * start)
        return (r[0::2] + 1j * r[1::2]).astype(dtype)
    z = splitmix64(seed, start, count)
    if dtype == np.float64:
        return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5
    if dtype == np.float32:
        return ((z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0) - np.float32(0.5)).astype(np.float32)
    raise TypeError(dtype)


def initial_block(n: int, k: int, seed: int, dtype=np.float64) -> np.ndarray:
    """X0 as an (n,k) Fortran-ordered array; element (i,j) uses counter j*n+i."""
    return splitmix_uniform(seed, n * k, dtype).reshape((n, k), order="F")


def laplacian_csr(grid, dtype=np.float64, potential=None):
    """CSR (rowptr int64, col int32, val) of the Dirichlet stencil Laplacian on grid=(gx[,gy[,gz]])."""
    g = tuple(int(v) for v in grid) + (1,) * (3 - len(grid))
    gx, gy, gz = g
    dim = len(grid)
    n = gx * gy * gz
    i = np.arange(n, dtype=np.int64)
    x = i % gx
    y = (i // gx) % gy
    z = i // (gx * gy)
    cols = []
    vals = []
    rows = []
    def add(mask, off, v):
        rows.append(i[mask]); cols.append(i[mask] + off); vals.append(np.full(mask.sum(), v))
    if gz > 1: add(z > 0, -gx * gy, -1.0)
    if gy > 1: add(y > 0, -gx, -1.0)
    add(x > 0, -1, -1.0)
    d = np.full(n, 2.0 * dim)
    if potential is not None:
        d = d + np.asarray(potential, dtype=np.float64)
    rows.append(i); cols.append(i); vals.append(d)
    add(x < gx - 1, 1, -1.0)
    if gy > 1: add(y < gy - 1, gx, -1.0)
    if gz > 1: add(z < gz - 1, gx * gy, -1.0)
    r = np.concatenate(rows); c = np.concatenate(cols); v = np.concatenate(vals)
    order = np.lexsort((c, r))
    r, c, v = r[order], c[order], v[order]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, c.astype(np.int32), v.astype(dtype)


def laplacian_eigs(grid, nev: int) -> np.ndarray:
    """Smallest nev analytic eigenvalues: sum_axes (2 - 2 cos(i pi/(g+1)))."""
    axes = [2.0 - 2.0 * np.cos(np.arange(1, g + 1) * np.pi / (g + 1)) for g in grid]
    lim = [min(len(a), max(8, int(4 * nev ** (1.0 / len(grid))) + 8)) for a in axes]
    tot = axes[0][: lim[0]]
    for a, l in zip(axes[1:], lim[1:]):
        tot = np.add.outer(tot, a[:l]).ravel()
    tot.sort()
    return tot[:nev]


def bdg_eigs(grid, nev: int, shift: float, d_abs: float) -> np.ndarray:
    """Positive-signature spectrum of A=[[K+s,d],[conj d,K+s]], B=diag(I,-I): sqrt((e+s)^2-|d|^2)."""
    e = laplacian_eigs(grid, nev)
    return np.sqrt((e + shift) ** 2 - d_abs ** 2)


def mass_diagonal(n: int, seed: int = 3, dtype=np.float64) -> np.ndarray:
    """SPD diagonal mass b_i = 0.5 + u_i, u uniform[0,1) (config C3)."""
    return (splitmix_uniform(seed, n, np.float64) + 1.0).astype(dtype)


def harmonic_potential(grid, omega: float = 0.05) -> np.ndarray:
    """v = 0.5*omega^2*r^2 about the grid centre (makes Jacobi T non-trivial)."""
    g = tuple(grid) + (1,) * (3 - len(grid))
    ax = [np.arange(m, dtype=np.float64) - 0.5 * (m - 1) for m in g]
    r2 = ax[0][None, None, :] ** 2 + ax[1][None, :, None] ** 2 + ax[2][:, None, None] ** 2
    return (0.5 * omega * omega * r2).ravel()
