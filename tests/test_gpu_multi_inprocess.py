"""The reference entry points on SEVERAL GPUs inside one call (csrc/multigpu.cu: lb2_set_num_gpus / LB2_GPUS): one process,
one worker thread per device, operators re-created as row slabs, halos through peer access, Gram sums through an in-process
NCCL communicator, X0 / eigenvectors moved by every device for its own rows.  Needs a box with >= 2 GPUs (gpurun --gpus 2);
skipped on a single-GPU box.  Each case is compared with the SAME call on one GPU."""
import numpy as np
import pytest

from lobpcg_b200 import api
from lobpcg_b200 import problems as pr

pytestmark = pytest.mark.gpu


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(ngpus() < 2, reason="needs >= 2 GPUs")


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b))))


def both(fn, n_multi):
    """fn() runs a reference-facing solve; returns (single-GPU result, multi-GPU result, gpus used)."""
    L = api.lib()
    L.lb2_set_num_gpus(1)
    r1 = fn()
    L.lb2_set_num_gpus(n_multi)
    try:
        r2 = fn()
        used = int(L.lb2_last_num_gpus())
    finally:
        L.lb2_set_num_gpus(0)
    return r1, r2, used


@needs2
def test_stencil_solve_on_all_gpus_of_the_box():
    g, nev, k = (48, 48, 48), 12, 24
    n = 48 ** 3
    X0 = pr.initial_block(n, k, 7)
    nd = min(ngpus(), 8)
    while 48 % nd:
        nd -= 1
    r1, r2, used = both(lambda: api.lobpcg(api.stencil_op(g, np.float64), X0, nev, 1e-8, 4000), nd)
    assert used == nd >= 2
    assert r1["status"] == r2["status"] == 0 and r1["converged"] == r2["converged"] == nev
    assert relerr(r2["eig"][:nev], r1["eig"][:nev]) < 1e-10
    assert relerr(r2["eig"][:nev], pr.laplacian_eigs(g, nev)) < 1e-10
    X = r2["X"]
    assert np.linalg.norm(X.T @ X - np.eye(k)) < 1e-8          # every device wrote its own rows of the eigenvectors


@needs2
def test_pencil_with_mass_and_polynomial_preconditioner_on_two_gpus():
    g, nev, k = (32, 32, 32), 8, 16
    n = 32 ** 3
    b = pr.mass_diagonal(n)
    X0 = pr.initial_block(n, k, 7)

    def run():
        A = api.stencil_op(g, np.float64, potential=pr.harmonic_potential(g, 0.2))
        return api.lobpcg(A, X0, nev, 1e-8, 2000, B=api.diag_op(b, np.float64), T=api.chebyshev_op(A, 8, 0.3, 0.0))
    r1, r2, used = both(run, 2)
    assert used == 2 and r1["converged"] == r2["converged"] == nev
    assert relerr(r2["eig"][:nev], r1["eig"][:nev]) < 1e-10
    X = r2["X"]
    assert np.linalg.norm(X.T @ (b[:, None] * X) - np.eye(k)) < 1e-8


@needs2
def test_general_csr_on_two_gpus(monkeypatch):
    g, nev, k = (24, 24, 24), 6, 12
    n = 24 ** 3
    monkeypatch.setenv("LB2_CSR_NO_STENCIL_DETECT", "1")
    rp, c, v = pr.laplacian_csr(g, potential=pr.harmonic_potential(g, 0.3))
    X0 = pr.initial_block(n, k, 7)
    r1, r2, used = both(lambda: api.lobpcg(api.csr_op(rp, c, v), X0, nev, 1e-8, 3000), 2)
    assert used == 2 and r1["converged"] == r2["converged"] == nev
    assert relerr(r2["eig"][:nev], r1["eig"][:nev]) < 1e-10


@needs2
def test_indefinite_bdg_on_two_gpus():
    g, nev, k = (16, 16, 16), 4, 8
    m = 16 ** 3
    bd = np.concatenate([np.ones(m), -np.ones(m)])
    X0 = pr.initial_block(2 * m, k, 13, np.complex128)
    X0[m:] *= 0.1
    shift, d = 0.5, 0.5 * np.exp(0.7j)

    def run():
        return api.lobpcg(api.bdg_op(g, np.complex128, shift, d), X0, nev, 1e-9, 4000, B=api.diag_op(bd, np.complex128),
                          indefinite=True)
    r1, r2, used = both(run, 2)
    assert used == 2 and r1["converged"] == r2["converged"] == nev
    assert relerr(r2["eig"][:nev], r1["eig"][:nev]) < 1e-10
    assert relerr(r2["eig"][:nev], pr.bdg_eigs(g, nev, shift, abs(d))) < 1e-10


@needs2
def test_operator_that_cannot_be_partitioned_runs_on_one_gpu():
    n, nev, k = 1000, 3, 6
    A = np.diag(np.arange(1.0, n + 1))
    r1, r2, used = both(lambda: api.lobpcg(api.dense_op(A), pr.initial_block(n, k, 5), nev, 1e-9, 2000), 2)
    assert used == 1 and r2["converged"] == nev
    assert relerr(r2["eig"][:nev], [1.0, 2.0, 3.0]) < 1e-9
