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
dist.shutdown(ctx)
