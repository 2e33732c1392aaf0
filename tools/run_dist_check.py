"""2+ GPU parity check of the row-partitioned solver: torchrun --nproc-per-node N tools/run_dist_check.py"""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import torch
from lobpcg_b200 import api, dist, problems as pr

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
ctx = api.Context(lr)
dist.init_process_group(ctx, "nccl")
g = (32, 32, 32); n = 32 ** 3; nev, k = 6, 12
part = dist.SlabPartition(*g, world, rank)
A = dist.partitioned_stencil(ctx, part, np.float64, k)
s = api.Solver(ctx, A, n, k, nev, np.float64, 1e-8, 5000, device_seed=7)
dist.attach(s, part)
s.init()
s.step(10 ** 6)
r = s.finish()
an = pr.laplacian_eigs(g, nev)
err = np.max(np.abs(r["eig"][:nev] - an) / an)
print(f"rank {rank}: iter {r['iter']} conv {r['converged']} max rel err vs analytic {err:.2e}", flush=True)
assert r["converged"] == nev and err < 1e-10
# eigenvectors: rows of this rank only were written
X = r["X"][part.rows()]
assert np.all(np.isfinite(X)) and np.abs(X).max() > 0
s.close()
# same problem with the built-in polynomial preconditioner: its inner stencil applies read halo planes of the
# workspace blocks (the non-current slab) from the neighbour's arena
T = api.chebyshev_op(A, 8, 0.3, 0.0)
s2 = api.Solver(ctx, A, n, k, nev, np.float64, 1e-8, 5000, T=T, device_seed=7)
dist.attach(s2, part)
s2.init()
s2.step(10 ** 6)
r2 = s2.finish()
err2 = np.max(np.abs(r2["eig"][:nev] - an) / an)
print(f"rank {rank}: chebyshev T: iter {r2['iter']} conv {r2['converged']} max rel err vs analytic {err2:.2e}", flush=True)
assert r2["converged"] == nev and err2 < 1e-10 and r2["iter"] * 3 < r["iter"]
s2.close()
# ... and evaluated in float inside the double solve (float blocks in the same arena region, float halo planes)
T3 = api.chebyshev_op(A, 8, 0.3, 0.0, mixed=True)
s3 = api.Solver(ctx, A, n, k, nev, np.float64, 1e-8, 5000, T=T3, device_seed=7)
dist.attach(s3, part)
s3.init()
s3.step(10 ** 6)
r3 = s3.finish()
err3 = np.max(np.abs(r3["eig"][:nev] - an) / an)
print(f"rank {rank}: mixed-precision chebyshev T: iter {r3['iter']} conv {r3['converged']} max rel err vs analytic {err3:.2e}", flush=True)
assert r3["converged"] == nev and err3 < 1e-10 and r3["iter"] * 3 < r["iter"]
# the same Laplacian with a harmonic trap as a row-partitioned CSR matrix (lb2_op_csr_slab: the kernel reads rows of the
# neighbouring blocks from the neighbours' arenas), Jacobi-free and with the polynomial preconditioner over the CSR inner
# operator (unfused Chebyshev steps); reference = the single-process stencil operator on the same problem, run on rank 0's
# GPU by every rank
pot = pr.harmonic_potential(g, 0.3)
rp, cc, vv = pr.laplacian_csr(g, potential=pot)
Ac = dist.partitioned_csr(part, rp, cc, vv)
hi = dist.global_spec_hi(Ac)
X0 = pr.initial_block(n, k, 7)
ref = api.lobpcg(api.stencil_op(g, np.float64, potential=pot), X0, nev, 1e-8, 5000)
for name, Tc in (("plain", None), ("chebyshev T", api.chebyshev_op(Ac, 8, 0.3, hi))):
    s4 = api.Solver(ctx, Ac, n, k, nev, np.float64, 1e-8, 5000, T=Tc, X0=X0)
    dist.attach(s4, part)
    s4.init()
    s4.step(10 ** 6)
    r4 = s4.finish()
    err4 = np.max(np.abs(r4["eig"][:nev] - ref["eig"][:nev]) / ref["eig"][:nev])
    print(f"rank {rank}: partitioned CSR ({name}): iter {r4['iter']} conv {r4['converged']} max rel err vs single-GPU "
          f"stencil solve {err4:.2e} (spectrum bound {hi:.3f})", flush=True)
    assert r4["converged"] == nev and err4 < 1e-10
    if Tc is not None:
        assert r4["iter"] * 3 < ref["iter"]
    else:
        # eigenvectors: this rank's rows, same as the single-GPU run up to sign
        Xl = r4["X"][part.rows(), :nev]
        assert np.all(np.isfinite(Xl))
    s4.close()
# config C4's shape on several GPUs: BdG pencil (A, B = diag(I, -I)) through z_ilobpcg's device path, both fields split by the
# same z-slabs (lb2_op_bdg_slab); analytic positive-signature spectrum, and the single-process solve on the same X0
gb = (16, 16, 16); mb = 16 ** 3; nevb, kb = 4, 8
shift, dcpl = 0.5, 0.5 * np.exp(0.7j)
partb = dist.SlabPartition(*gb, world, rank)
X0b = pr.initial_block(2 * mb, kb, 13, np.complex128); X0b[mb:] *= 0.1      # B-positive start
bdiag = np.concatenate([np.ones(mb), -np.ones(mb)])
refb = api.lobpcg(api.bdg_op(gb, np.complex128, shift, dcpl), X0b, nevb, 1e-9, 3000, B=api.diag_op(bdiag, np.complex128),
                  indefinite=True)
Ab = dist.partitioned_bdg(partb, np.complex128, shift, dcpl)
rows_b = dist.bdg_local_rows(partb)
Bb = api.diag_op(bdiag[rows_b], np.complex128)
s5 = api.Solver(ctx, Ab, 2 * mb, kb, nevb, np.complex128, 1e-9, 3000, B=Bb, X0=X0b, indefinite=True)
dist.attach(s5, partb)
s5.init()
s5.step(10 ** 6)
r5 = s5.finish()
anb = pr.bdg_eigs(gb, nevb, shift, abs(dcpl))
err5 = np.max(np.abs(r5["eig"][:nevb] - anb) / anb)
err5r = np.max(np.abs(r5["eig"][:nevb] - refb["eig"][:nevb]) / refb["eig"][:nevb])
print(f"rank {rank}: partitioned BdG pencil (ilobpcg): iter {r5['iter']} (single GPU {refb['iter']}) conv {r5['converged']} "
      f"max rel err vs analytic {err5:.2e}, vs single-GPU solve {err5r:.2e}, signatures {r5['sig'][:nevb]}", flush=True)
assert r5["converged"] == refb["converged"] == nevb and err5 < 1e-10 and err5r < 1e-10
assert np.all(r5["sig"][:nevb] == 1)
Xb = r5["X"][rows_b, :nevb]
assert np.all(np.isfinite(Xb)) and np.abs(Xb).max() > 0
s5.close()
dist.shutdown(ctx)
