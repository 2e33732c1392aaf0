"""Time-to-solution of a BASELINE config on 1..N GPUs (torchrun for N>1):
   python tools/full_solve.py G NEV [tol] [maxiter] [cheb=DEGREE,LO]   -> JSON line with iterations, seconds, eigenvalue errors
cheb=DEGREE,LO[,mixed] uses the built-in polynomial preconditioner T = lb2_op_chebyshev(A, DEGREE, LO, Gershgorin bound)."""
import os, sys, time, json
sys.path.insert(0, ".")
import numpy as np
import torch
from lobpcg_b200 import api, dist, problems as pr

cheb = next((a.split("=")[1] for a in sys.argv if a.startswith("cheb=")), None)
argv = [a for a in sys.argv if not a.startswith("cheb=")]
g = int(argv[1]); nev = int(argv[2]); tol = float(argv[3]) if len(argv) > 3 else 1e-8
maxit = int(argv[4]) if len(argv) > 4 else 5000
k = 2 * nev; n = g ** 3
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
ctx = api.Context(lr)
if world > 1:
    dist.init_process_group(ctx, "nccl")
    part = dist.SlabPartition(g, g, g, world, rank)
    A = dist.partitioned_stencil(ctx, part, np.float64, k)
else:
    part = None
    A = api.stencil_op((g, g, g), np.float64)
T = None
if cheb:
    deg, lo = cheb.split(",")[:2]
    T = api.chebyshev_op(A, int(deg), float(lo), 0.0, mixed=cheb.endswith(",mixed"))
s = api.Solver(ctx, A, n, k, nev, np.float64, tol, maxit, T=T, device_seed=7)
if part is not None:
    dist.attach(s, part)
t0 = time.time(); s.init(); ctx.sync(); t_init = time.time() - t0
s.reset_stats()
t0 = time.time()
hist = []
while True:
    done = s.step(25)
    p = s.progress(); hist.append((p["iter"], p["converged"], p["use_ortho"], round(time.time() - t0, 2)))
    if rank == 0:
        print("progress", hist[-1], flush=True)
    if done < 25:
        break
ctx.sync(); t_solve = time.time() - t0
p = s.progress()
import ctypes as C
st = s.state_
eig = st.eigvals() if False else None
# eigenvalues/residuals without downloading X: they are host-side after every pass
lib = api.lib()
eigs, resn = s.results()
stats = s.stats()
if rank == 0:
    out = dict(grid=g, nev=nev, k=k, n_gpus=world, tol=tol, preconditioner=("chebyshev degree,lo=" + cheb) if cheb else None, init_s=t_init, solve_s=t_solve, passes=p["iter"] + 1,
               converged=p["converged"], use_ortho=p["use_ortho"], s_per_pass=t_solve / (p["iter"] + 1),
               phases_ms_per_pass={kk: round(v["ms"] / (p["iter"] + 1), 2) for kk, v in stats.items()})
    an = pr.laplacian_eigs((g, g, g), nev)
    out["max_rel_eig_err_vs_analytic"] = float(np.max(np.abs(eigs[:nev] - an) / an))
    out["max_resnorm"] = float(resn[:nev].max())
    print(json.dumps(out), flush=True)
if world > 1:
    dist.shutdown(ctx)
