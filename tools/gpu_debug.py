"""Quick GPU sanity driver used while developing (python tools/gpu_debug.py)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from lobpcg_b200 import api, problems as pr

ctx = api.Context()
rng = np.random.default_rng(0)

def chk(name, got, ref, tol):
    err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)
    print(f"{name:40s} relerr={err:.3e} {'OK' if err < tol else 'FAIL'}", flush=True)

for dt, tol in [(np.float64, 1e-12), (np.float32, 1e-4), (np.complex128, 1e-12), (np.complex64, 1e-4)]:
    for (n, ma, mb) in [(5000, 7, 5), (20001, 60, 60), (33333, 130, 70)]:
        A = rng.standard_normal((n, ma)).astype(dt); B = rng.standard_normal((n, mb)).astype(dt)
        if np.dtype(dt).kind == 'c':
            A = A + 1j * rng.standard_normal((n, ma)).astype(dt); B = B + 1j * rng.standard_normal((n, mb)).astype(dt)
        dA = api.DeviceArray.from_numpy(ctx, A); dB = api.DeviceArray.from_numpy(ctx, B)
        G = api.gram(ctx, dA, dB).numpy(ctx)
        chk(f"gram {np.dtype(dt).name} {n}x{ma}x{mb}", G, A.conj().T @ B, tol)
        if ma == mb or True:
            G2 = api.gram(ctx, dA, dA, upper=True).numpy(ctx)
            chk(f"gram-upper {np.dtype(dt).name} {n}x{ma}", G2, A.conj().T @ A, tol)
        Cm = rng.standard_normal((ma, mb)).astype(dt)
        dC = api.DeviceArray.from_numpy(ctx, Cm)
        O0 = rng.standard_normal((n, mb)).astype(dt)
        dO = api.DeviceArray.from_numpy(ctx, O0)
        api.tall_nn(ctx, dA, dC, dO, alpha=-1.0, beta=1.0)
        chk(f"tall_nn {np.dtype(dt).name} {n}x{ma}x{mb}", dO.numpy(ctx), O0 - A @ Cm, tol)

# spmm
for dt, tol in [(np.float64, 1e-13), (np.complex128, 1e-13), (np.float32, 1e-5)]:
    for grid in [(100,), (37, 21), (33, 18, 21), (64, 64, 64)]:
        n = int(np.prod(grid)); nc = 7
        X = rng.standard_normal((n, nc)).astype(dt)
        rp, c, v = pr.laplacian_csr(grid)
        import scipy.sparse as sp
        M = sp.csr_matrix((v, c, rp), shape=(n, n))
        ref = M @ X
        dX = api.DeviceArray.from_numpy(ctx, X)
        op = api.stencil_op(grid, dt)
        chk(f"stencil {np.dtype(dt).name} {grid}", op.apply(ctx, dX).numpy(ctx), ref, tol)
        opc = api.csr_op(rp, c, v.astype(dt))
        chk(f"csr {np.dtype(dt).name} {grid}", opc.apply(ctx, dX).numpy(ctx), ref, tol)

# residual
n, nc = 30011, 9
AX = rng.standard_normal((n, nc)); BX = rng.standard_normal((n, nc)); lam = rng.standard_normal(nc)
W, ss = api.residual(ctx, api.DeviceArray.from_numpy(ctx, AX), api.DeviceArray.from_numpy(ctx, BX), api.DeviceArray.from_numpy(ctx, lam))
chk("residual W", W.numpy(ctx), AX - BX * lam, 1e-14)
chk("residual sumsq", ss.numpy(ctx), ((AX - BX * lam) ** 2).sum(0), 1e-13)
X0 = api.fill_uniform(ctx, 1000, 3, np.float64, 7).numpy(ctx)
print("fill_uniform bit-exact:", np.array_equal(X0, pr.initial_block(1000, 3, 7)))

# solver C1
g = (100, 100); n = 10000; nev = 10; k = 20
A = api.stencil_op(g, np.float64)
X0 = pr.initial_block(n, k, 7)
t = time.time()
r = api.lobpcg(A, X0, nev, 1e-8, 5000)
an = pr.laplacian_eigs(g, nev)
print("C1:", time.time() - t, "s iter", r['iter'], "conv", r['converged'], "max rel err vs analytic", np.max(np.abs(r['eig'][:nev] - an) / an), "res max", r['res'][:nev].max())
