"""Generates tests/golden/*.json|npz.  Run HERE (container with /root/reference and oracle/_ref built):

    make -C oracle && python tests/golden/make_golden.py

(1) known_answers.json — the reference's own known-answer vectors for the hot path, restated with the
    file:line they come from.
(2) reference_runs.npz  — outputs of the UNMODIFIED reference (oracle/_ref) on seeded inputs produced by
    lobpcg_b200.problems: full solver runs and single calls of d_svqb / d_ortho_drop /
    d_rayleigh_ritz / d_rayleigh_ritz_modified / d_gram_* / d_get_residual.
The GPU box has no /root/reference; tests there read only these files.
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sp  # noqa: F401  (import before oracle/_ref is dlopen'ed: importing scipy afterwards crashes)

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from lobpcg_b200 import problems as pr  # noqa: E402
from oracle import ref_bindings as rb  # noqa: E402

OUT = Path(__file__).resolve().parent


def known_answers():
    ka = {
        "_source": "pstuermer/LOBPCG tests/, restated by hand; indices are column-major",
        "gram_self_d_general": {  # tests/test_gram.c:156-174
            "U": [[1, 0], [0, 1], [1, 1]], "G_upper": {"00": 2.0, "01": 1.0, "11": 2.0}, "tol": 1e-12},
        "gram_self_z_with_B": {  # tests/test_gram.c:203-227
            "U_re": [[1, 0], [0, 1], [0, 0]], "U_im": [[1, 0], [0, -1], [0, 0]], "Bdiag": [2.0, 3.0, 4.0],
            "G_upper": {"00": 4.0, "01": 0.0, "11": 6.0}, "tol": 1e-12},
        "residual_noneigvec_real": {  # tests/test_residual.c:277-300
            "Adiag": [1.0, 2.0, 3.0], "x": [1.0, 2.0, 3.0], "lambda": 2.0, "R": [-1.0, 0.0, 3.0]},
        "residual_with_B_real": {  # tests/test_residual.c (with-B case): A=diag(1,2,3), B=diag(2,2,2)... R=[-5,0,7]
            "note": "R = A x - lambda B x with A=diag(1,4,9)?? not restated; covered by reference_runs"},
        "rayleigh_ritz_4x4": {  # tests/test_rayleigh_ritz.c:53-73
            "A": [4, 1, 2, 0, 1, 3, 0, 1, 2, 0, 5, 2, 0, 1, 2, 6],
            "S": [1, -1, 1, -1, 1, 1, -1, -2],
            "eig": [4.270248, 5.507529], "eig_tol": 1e-4,
            "Xnew": [0.326799989, -0.658521176, 0.658521176, -0.160939395,
                     0.475957037, 0.218703777, -0.218703777, -0.823287444], "X_tol": 1e-6},
        "lobpcg_dense_4x4": {  # tests/test_lobpcg.c:87-92,105-108
            "A": [4, 1, 2, 0, 1, 3, 0, 1, 2, 0, 5, 2, 0, 1, 2, 6],
            "eig": [1.338399579631295e+00, 3.463077212970466e+00, 5.0, 8.198523207398235e+00]},
        "lobpcg_dense_6x6": {  # tests/test_lobpcg.c:94-114
            "A": [4.0, 1.0, 2.0, 0.0, 1.0, 0.5, 1.0, 3.0, 0.0, 1.0, 0.5, 0.0, 2.0, 0.0, 5.0, 2.0, 1.0, 1.0,
                  0.0, 1.0, 2.0, 6.0, 1.5, 0.0, 1.0, 0.5, 1.0, 1.5, 5.0, 2.0, 0.5, 0.0, 1.0, 0.0, 2.0, 4.0],
            "eig": [1.208742643127633e+00, 2.230197331224639e+00, 3.615464945758393e+00,
                    4.717703764957660e+00, 5.517221003524097e+00, 9.710670311407574e+00]},
        "lobpcg_softlock_diag30": {  # tests/test_lobpcg.c:455-500
            "n": 30, "nev": 3, "sizeSub": 6, "tol": 1e-10, "eig": [1.0, 2.0, 3.0], "eig_tol": 1e-8},
        "lobpcg_laplacian_1d": {  # tests/test_lobpcg.c:349-393: n=100, nev=3, sizeSub=5, tol 1e-4, (k pi)^2 within 1%
            "n": 100, "nev": 3, "sizeSub": 5, "tol": 1e-4, "rel_tol": 1e-2},
    }
    del ka["residual_with_B_real"]
    (OUT / "known_answers.json").write_text(json.dumps(ka, indent=1))


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def single_calls(out):
    """Direct calls of the reference's L2-L4 helpers (lobpcg.h:98-555) on seeded inputs."""
    L = rb.lib()
    u64, dbl = C.c_uint64, C.c_double
    n, nu, nv = 500, 6, 9
    U = pr.initial_block(n, nu, 11)
    V, _ = np.linalg.qr(pr.initial_block(n, nv, 12))
    V = np.asfortranarray(V)
    bdiag = pr.mass_diagonal(n, seed=5)
    B = rb.op_diag(bdiag, np.float64)
    out["sc_U"], out["sc_V"], out["sc_bdiag"] = U, V, bdiag
    for tag, Bh in (("I", None), ("B", B.handle)):
        # gram_self / gram_cross (src/gram/gram_impl.inc:49,85)
        G = np.zeros((nu, nu), order="F"); wrk = np.zeros((n, max(nu, nv)), order="F")
        L.d_gram_self.argtypes = [C.c_void_p, u64, u64, C.c_void_p, C.c_void_p, u64, C.c_void_p]
        L.d_gram_self(vp(U), n, nu, Bh, vp(G), nu, vp(wrk))
        out[f"sc_gram_self_{tag}"] = G.copy()
        Gc = np.zeros((nv, nu), order="F")
        L.d_gram_cross.argtypes = [C.c_void_p, u64, C.c_void_p, u64, u64, C.c_void_p, C.c_void_p, u64, C.c_void_p]
        L.d_gram_cross(vp(V), nv, vp(U), nu, n, Bh, vp(Gc), nv, vp(wrk))
        out[f"sc_gram_cross_{tag}"] = Gc.copy()
        # svqb (src/ortho/svqb_impl.inc:48)
        U2 = U.copy(order="F"); U2[:, 3] = U2[:, 1]  # duplicate column => one drop with drop='y'
        w1 = np.zeros((nu, nu), order="F"); w2 = np.zeros((n, nu), order="F"); w3 = np.zeros((n, nu), order="F")
        L.d_svqb.restype = u64
        L.d_svqb.argtypes = [u64, u64, dbl, C.c_char, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        nret = L.d_svqb(n, nu, 1e-12, b"y", vp(U2), vp(w1), vp(w2), vp(w3), Bh)
        out[f"sc_svqb_{tag}_nret"] = np.array(nret)
        out[f"sc_svqb_{tag}_U"] = U2[:, :nret].copy()
        # ortho_drop (src/ortho/ortho_drop_impl.inc:43); V made B-orthonormal first through the reference itself
        V2 = V.copy(order="F")
        w1 = np.zeros((n, nv), order="F"); w2 = np.zeros((n, nv), order="F"); w3 = np.zeros((n, nv), order="F")
        L.d_svqb(n, nv, 1e-12, b"n", vp(V2), vp(w1), vp(w2), vp(w3), Bh)
        U3 = U.copy(order="F")
        L.d_ortho_drop.restype = u64
        L.d_ortho_drop.argtypes = [u64, u64, u64, dbl, dbl] + [C.c_void_p] * 6
        w1 = np.zeros((n, nv), order="F"); w2 = np.zeros((n, nv), order="F"); w3 = np.zeros((n, nv), order="F")
        nret = L.d_ortho_drop(n, nu, nv, 1e-12, 1e-12, vp(U3), vp(V2), vp(w1), vp(w2), vp(w3), Bh)
        out[f"sc_ortho_{tag}_V"] = V2
        out[f"sc_ortho_{tag}_nret"] = np.array(nret)
        out[f"sc_ortho_{tag}_U"] = U3[:, :nret].copy()
    # get_residual (src/residual/residual_impl.inc:32)
    A = rb.op_stencil((n,), np.float64)
    lam = np.linspace(0.5, 2.0, nu)
    W = np.zeros((n, nu), order="F"); wrk = np.zeros((n, nu), order="F")
    L.d_get_residual.argtypes = [u64, u64] + [C.c_void_p] * 7
    L.d_get_residual(n, nu, vp(U), None, vp(W), vp(lam), vp(wrk), A.handle, B.handle)
    out["sc_residual_lam"], out["sc_residual_W"] = lam, W.copy()
    # rayleigh_ritz (src/rayleigh/rayleigh_ritz_impl.inc:37)
    Cx = np.zeros((nu, nu), order="F"); ev = np.zeros(nu); D = np.zeros(3 * nu)
    w1 = np.zeros((n, 2 * nu), order="F"); w2 = np.zeros((n, 3 * nu), order="F"); w3 = np.zeros((n, nu), order="F")
    L.d_rayleigh_ritz.argtypes = [u64, u64] + [C.c_void_p] * 9
    L.d_rayleigh_ritz(n, nu, vp(U), vp(Cx), vp(ev), vp(w1), vp(w2), vp(w3), vp(D), A.handle, B.handle)
    out["sc_rr_eig"], out["sc_rr_X"] = ev.copy(), (U @ Cx).copy()
    # rayleigh_ritz_modified, Cholesky branch, mult=3 (src/rayleigh/rayleigh_ritz_modified_impl.inc:42)
    nx = 4
    S = pr.initial_block(n, 3 * nx, 21)
    AX = A.apply(S[:, :nx])
    Cx = np.zeros((3 * nx, nx), order="F"); Cp = np.zeros((3 * nx, nx), order="F"); ev = np.zeros(nx)
    rre = np.zeros(3 * nx); tau = np.zeros(3 * nx); D = np.zeros(3 * nx)
    w1 = np.zeros((n, 2 * nx), order="F"); w2 = np.zeros((n, 3 * nx), order="F"); w3 = np.zeros((n, nx), order="F")
    uo = C.c_uint8(0)
    L.d_rayleigh_ritz_modified.argtypes = [u64, u64, u64, u64, u64, C.POINTER(C.c_uint8)] + [C.c_void_p] * 13
    L.d_rayleigh_ritz_modified(n, nx, 3, 0, nx, C.byref(uo), vp(S), vp(AX), vp(w1), vp(w2), vp(w3), vp(Cx), vp(Cp),
                               vp(ev), vp(rre), vp(tau), vp(D), A.handle, B.handle)
    out["sc_rrm_S"], out["sc_rrm_eig"], out["sc_rrm_useortho"] = S, ev.copy(), np.array(uo.value)
    out["sc_rrm_X"], out["sc_rrm_P"] = (S @ Cx).copy(), (S @ Cp).copy()


def solver_runs(out):
    def run(tag, A, X0, nev, tol, maxit, B=None, T=None):
        r = rb.solve(A, X0, nev, tol, maxit, B=B, T=T)
        out[f"run_{tag}_eig"] = r["eig"]
        out[f"run_{tag}_res"] = r["res"]
        out[f"run_{tag}_meta"] = np.array([r["iter"], r["converged"], nev, X0.shape[1]])
        print(tag, "iter", r["iter"], "conv", r["converged"], r["eig"][:nev])

    # C1: 2-D 100x100, nev 10, k 20, tol 1e-8, seed 7
    run("c1", rb.op_stencil((100, 100), np.float64), pr.initial_block(10000, 20, 7), 10, 1e-8, 5000)
    # 1-D n=100 (shape of tests/test_lobpcg.c:349-393, unscaled stencil, k = 2 nev)
    run("lap1d", rb.op_stencil((100,), np.float64), pr.initial_block(100, 6, 123), 3, 1e-8, 5000)
    # generalized pencil + Jacobi T (C3 shape at 12^3)
    n = 12 ** 3
    b = pr.mass_diagonal(n)
    run("gen3d", rb.op_stencil((12, 12, 12), np.float64), pr.initial_block(n, 8, 7), 4, 1e-8, 3000,
        B=rb.op_diag(b, np.float64), T=rb.op_diag(np.full(n, 1 / 6.0), np.float64))
    # CSR + harmonic potential + non-trivial Jacobi (C2 variant at 16^3)
    g = (16, 16, 16); n = 16 ** 3
    pot = pr.harmonic_potential(g, 0.3)
    rp, c, v = pr.laplacian_csr(g, potential=pot)
    run("pot3d", rb.op_csr(rp, c, v), pr.initial_block(n, 12, 7), 6, 1e-8, 3000,
        T=rb.op_diag(1.0 / (6.0 + pot), np.float64))
    # float (C3-f32 shape at 12^3)
    n = 12 ** 3
    run("gen3d_f32", rb.op_stencil((12, 12, 12), np.float32), pr.initial_block(n, 8, 7, np.float32), 4, 1e-4, 3000,
        B=rb.op_diag(b, np.float32))
    # soft-locking: diag(1..30), nev 3, k 6 (tests/test_lobpcg.c:455-500)
    run("softlock", rb.op_diag(np.arange(1.0, 31.0), np.float64), pr.initial_block(30, 6, 5), 3, 1e-10, 500)
    # complex Hermitian definite: real stencil in complex arithmetic, 10^3
    n = 1000
    run("z3d", rb.op_stencil((10, 10, 10), np.complex128), pr.initial_block(n, 8, 9, np.complex128), 4, 1e-8, 3000)


def preconditioned_cases():
    """Runs with the polynomial preconditioner T = p(A) (lb2_op_chebyshev / ref_<p>_op_cheb): (A builder args, T args)."""
    g = (16, 16, 16)
    pot = pr.harmonic_potential(g, 0.3)
    return {
        "cheb3d": dict(grid=g, pot=None, nev=6, k=12, degree=8, lo=0.3, hi=12.0, mass=False, csr=False),
        "cheb_gen3d": dict(grid=(12, 12, 12), pot=None, nev=4, k=8, degree=5, lo=0.5, hi=12.0, mass=True, csr=False),
        "cheb_pot_csr": dict(grid=g, pot=pot, nev=6, k=12, degree=6, lo=0.4, hi=12.0 + float(pot.max()), mass=False, csr=True),
    }


def preconditioned_runs(out):
    for tag, c in preconditioned_cases().items():
        g = c["grid"]; n = g[0] * g[1] * g[2]
        if c["csr"]:
            rp, cc, v = pr.laplacian_csr(g, potential=c["pot"])
            A = rb.op_csr(rp, cc, v)
        else:
            A = rb.op_stencil(g, np.float64, potential=c["pot"])
        B = rb.op_diag(pr.mass_diagonal(n), np.float64) if c["mass"] else None
        T = rb.op_cheb(A, c["degree"], c["lo"], c["hi"])
        X0 = pr.initial_block(n, c["k"], 7)
        r = rb.solve(A, X0, c["nev"], 1e-8, 3000, B=B, T=T)
        r0 = rb.solve(A, X0, c["nev"], 1e-8, 3000, B=B)
        out[f"run_{tag}_eig"] = r["eig"]
        out[f"run_{tag}_res"] = r["res"]
        out[f"run_{tag}_meta"] = np.array([r["iter"], r["converged"], c["nev"], c["k"]])
        out[f"run_{tag}_iter_without_T"] = np.array([r0["iter"]])
        print(tag, "iter", r["iter"], "(without T:", r0["iter"], ") conv", r["converged"], r["eig"][:c["nev"]])


def indefinite_cases():
    """Inputs of the ilobpcg parity cases (shared with tests/test_gpu_solver.py through this module)."""
    cases = {}
    # BdG-style Hermitian pencil, config C4 at 6^3: A=[[K+c,d],[conj d,K+c]], B=diag(I,-I)
    g = (6, 6, 6); m = 216
    X0 = pr.initial_block(2 * m, 8, 13, np.complex128); X0[m:] *= 0.1      # B-positive start
    cases["ilob_bdg_z"] = dict(kind="bdg", grid=g, dtype=np.complex128, shift=0.5, d=0.5 * np.exp(0.7j),
                               bdiag=np.concatenate([np.ones(m), -np.ones(m)]), X0=X0, nev=4, tol=1e-9, it=3000)
    X0 = pr.initial_block(2 * m, 8, 14, np.float64); X0[m:] *= 0.1
    cases["ilob_bdg_d"] = dict(kind="bdg", grid=g, dtype=np.float64, shift=0.5, d=0.3,
                               bdiag=np.concatenate([np.ones(m), -np.ones(m)]), X0=X0, nev=4, tol=1e-9, it=3000)
    # block Laplacian / block swap (reference tests/test_ilobpcg.c:160-223): A=blkdiag(K,K), B=[[0,I],[I,0]], start [u;u]
    mm = 50
    rp, cc, vv = pr.laplacian_csr((mm,))
    K = sp.csr_matrix((vv, cc, rp), shape=(mm, mm))
    Ab = sp.block_diag([K, K]).tocsr(); Ab.sort_indices()
    Bb = sp.bmat([[None, sp.eye(mm)], [sp.eye(mm), None]]).tocsr(); Bb.sort_indices()
    u = pr.initial_block(mm, 6, 3)
    cases["ilob_swap_d"] = dict(kind="csr", A=(Ab.indptr, Ab.indices, Ab.data.astype(np.float64)),
                                B=(Bb.indptr, Bb.indices, Bb.data.astype(np.float64)),
                                X0=np.asfortranarray(np.vstack([u, u])), nev=3, tol=1e-8, it=3000)
    return cases


def indefinite_runs(out):
    for tag, c in indefinite_cases().items():
        if c["kind"] == "bdg":
            A = rb.op_bdg(c["grid"], c["dtype"], c["shift"], c["d"])
            B = rb.op_diag(c["bdiag"], c["dtype"])
        else:
            A, B = rb.op_csr(*c["A"]), rb.op_csr(*c["B"])
        r = rb.solve(A, c["X0"], c["nev"], c["tol"], c["it"], B=B, indefinite=True)
        out[f"run_{tag}_eig"] = r["eig"]
        out[f"run_{tag}_res"] = r["res"]
        out[f"run_{tag}_sig"] = r["sig"]
        out[f"run_{tag}_meta"] = np.array([r["iter"], r["converged"], c["nev"], c["X0"].shape[1]])
        print(tag, "iter", r["iter"], "conv", r["converged"], r["eig"][:c["nev"]], r["sig"][:c["nev"]])


if __name__ == "__main__":
    if not rb.available():
        raise SystemExit("build oracle/_ref first: make -C oracle")
    known_answers()
    out = {}
    single_calls(out)
    solver_runs(out)
    indefinite_runs(out)
    preconditioned_runs(out)
    old = OUT / "reference_runs.npz"
    if old.exists():   # fixtures are deterministic (fixed time() in the harness): regenerating must not change them
        prev = np.load(old)
        changed = [k for k in prev.files if k in out and not np.array_equal(prev[k], out[k])]
        print("keys changed by regeneration:", changed)
    np.savez_compressed(OUT / "reference_runs.npz", **out)
    print("wrote", OUT / "reference_runs.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB", flush=True)
    import os
    os._exit(0)  # skip interpreter teardown (operator handles would be freed after the library is gone)
