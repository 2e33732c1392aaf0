"""Full-size runs of the BASELINE configs C1..C4 on one B200 (C5 is bench.py / tools/full_solve.py):
   python tools/run_configs.py [C1 C2 C2csr C3d C3s C4] -> one JSON line per config with passes, time-to-solution,
   eigenvalue error vs the analytic spectrum (where one exists), residual norms and per-phase kernel times.

Inputs are exactly SURVEY.md §8(d): seeds, tolerances, k = 2*nev.  Everything goes through the resumable solver
handle of the C ABI (lb2_solver_*), X0 generated on the device with the portable splitmix64 generator."""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from lobpcg_b200 import api, problems as pr

CHEB = next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("cheb=")), None)   # cheb=DEGREE,LO: built-in T = p(A)
which = [a for a in sys.argv[1:] if not a.startswith("cheb=")] or ["C1", "C2", "C2csr", "C3d", "C3s", "C4"]


def precond(A):
    if not CHEB:
        return None
    deg, lo = CHEB.split(",")
    return api.chebyshev_op(A, int(deg), float(lo), 0.0)
ctx = api.Context(0)


def solve(name, A, n, nev, dtype, tol, B=None, T=None, indefinite=False, X0=None, maxit=20000, extra=None):
    k = 2 * nev
    s = api.Solver(ctx, A, n, k, nev, dtype, tol, maxit, B=B, T=T, X0=X0, device_seed=None if X0 is not None else 7,
                   indefinite=indefinite)
    t0 = time.time(); s.init(); ctx.sync(); t_init = time.time() - t0
    s.reset_stats()
    t0 = time.time()
    while s.step(50) == 50:
        pass
    ctx.sync(); t_solve = time.time() - t0
    p = s.progress()
    eigs, resn = s.results()
    passes = p["iter"] + 1
    out = dict(config=name + (f" + chebyshev T ({CHEB})" if CHEB and T is not None else ""), n=n, nev=nev, k=k, dtype=str(np.dtype(dtype)), tol=tol, init_s=round(t_init, 3),
               solve_s=round(t_solve, 3), passes=passes, iters_per_s=round(passes / t_solve, 3), converged=p["converged"],
               use_ortho=p["use_ortho"], max_resnorm=float(resn[:nev].max()),
               phases_ms_per_pass={kk: round(v["ms"] / passes, 3) for kk, v in s.stats().items()})
    if extra:
        out.update(extra(eigs[:nev]))
    s.close()
    print(json.dumps(out), flush=True)
    return eigs[:nev]


def err_vs(an):
    return lambda e: {"max_rel_eig_err_vs_analytic": float(np.max(np.abs(e - an) / np.abs(an)))}


if "C1" in which:   # 2-D 5-point 100x100, nev=10, unpreconditioned
    g = (100, 100)
    A = api.stencil_op(g, np.float64)
    solve("C1", A, 10000, 10, np.float64, 1e-8, T=precond(A), extra=err_vs(pr.laplacian_eigs(g, 10)))

for name in ("C2", "C2csr"):   # 3-D 7-point 128^3 as CSR, nev=64, Jacobi T
    if name not in which:
        continue
    g = (128, 128, 128)
    n = 128 ** 3
    if name == "C2csr":
        os.environ["LB2_CSR_NO_STENCIL_DETECT"] = "1"   # force the general CSR kernel
    rp, col, val = pr.laplacian_csr(g)
    A = api.csr_op(rp, col, val)
    os.environ.pop("LB2_CSR_NO_STENCIL_DETECT", None)
    T = precond(A) or api.diag_op(np.full(n, 1.0 / 6.0), np.float64)
    solve(name + (" (general CSR kernel)" if name == "C2csr" else " (CSR recognised as a stencil)"), A, n, 64, np.float64,
          1e-8, T=T, extra=err_vs(pr.laplacian_eigs(g, 64)))
    del A, rp, col, val

c3 = {}
for name, dt, tol in (("C3d", np.float64, 1e-8), ("C3s", np.float32, 1e-4)):   # pencil A x = lambda B x, 160^3, nev=100
    if name not in which:
        continue
    g = (160, 160, 160)
    n = 160 ** 3
    b = pr.mass_diagonal(n)
    A = api.stencil_op(g, dt)
    c3[name] = solve(name, A, n, 100, dt, tol, B=api.diag_op(b, dt), T=precond(A))
if len(c3) == 2:
    print(json.dumps({"config": "C3 float vs double", "max_rel_eig_diff": float(np.max(np.abs(c3["C3s"] - c3["C3d"]) /
                                                                                   np.abs(c3["C3d"])))}), flush=True)

if "C4" in which:   # indefinite LOBPCG on the BdG-style pencil, n = 2 * 80^3, nev=50, complex double
    g = (80, 80, 80)
    m = 80 ** 3
    n = 2 * m
    shift, d = 0.5, 0.5 * np.exp(0.7j)
    A = api.bdg_op(g, np.complex128, shift, d)
    Bd = np.concatenate([np.ones(m), -np.ones(m)])
    k = 100
    X0 = pr.initial_block(n, k, 7, np.complex128)
    X0[m:, :] *= 0.1   # B-positive start
    an = pr.bdg_eigs(g, 50, shift, abs(d))   # omega = sqrt((eps + c)^2 - |d|^2) = sqrt(eps (eps + 2c)) for |d| = c
    solve("C4", A, n, 50, np.complex128, 1e-8, B=api.diag_op(Bd, np.complex128), indefinite=True, X0=X0, extra=err_vs(an))
