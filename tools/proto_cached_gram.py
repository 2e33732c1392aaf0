"""CPU prototype (numpy) of the cached-Gram-block LOBPCG pass used to choose the refresh policy before
the CUDA implementation: per pass only the W columns of S^H B S and S^H A S are contracted over n; the [X P] blocks follow
from C^H G C on the small matrices.  Tracks the TRUE error of the cached blocks and the growth bound the solver uses.

    python tools/proto_cached_gram.py
"""
import sys
sys.path.insert(0, ".")
import numpy as np
import scipy.linalg as sla
from oracle import numpy_oracle as no
from lobpcg_b200 import problems as pr

H = lambda M: M.conj().T


def lobpcg_cached(A, X0, nev, tol, max_iter, B=None, T=None, refresh_tol=1e-12, log=None, period=64, policy='diag'):
    X = np.array(X0, copy=True)
    n, k = X.shape
    dtype = X.dtype
    eps = no._eps_tol(dtype)
    from lobpcg_b200.problems import splitmix_uniform
    rng = lambda cnt, dt: splitmix_uniform(12345, cnt, dt)
    anorm = no.estimate_norm(A, n, dtype, rng)
    bnorm = no.estimate_norm(B, n, dtype, rng) if B is not None else 1.0
    Cx, eig = no.rayleigh_ritz(X, A, B)
    X = X @ Cx
    AX = A(X)
    W = no.get_residual(X, AX, eig, A, B)
    use_ortho = 0
    converged = 0
    P = np.zeros((n, 0), dtype)
    it = 0
    Gb_c = Ga_c = None     # cached [X P] blocks
    bound = 0.0            # tracked bound of the cache error (relative)
    nrefresh = 0
    diag_dev = 0.0
    since = 0
    Bop = (lambda Y: Y) if B is None else B
    while it < max_iter:
        nconv = converged
        Pa = P[:, nconv:] if it else np.zeros((n, 0), dtype)
        Wa = W[:, nconv:] if it else W
        if T is not None:
            Wa = T(Wa)
        V = np.concatenate([X, Pa], axis=1)
        if use_ortho:
            Wa, _ = no.ortho_drop(Wa, V, eps, eps, B)
        mxp = V.shape[1]

        def build(Wa):
            S = np.concatenate([V, Wa], axis=1)
            AW = A(Wa)
            m = S.shape[1]
            Gb = np.empty((m, m), dtype); Ga = np.empty((m, m), dtype)
            # W columns from tall contractions
            Gb[:, mxp:] = H(S) @ Bop(Wa)
            Ga[:, mxp:] = H(S) @ AW
            Gb[mxp:, :mxp] = H(Gb[:mxp, mxp:]); Ga[mxp:, :mxp] = H(Ga[:mxp, mxp:])
            Gb[mxp:, mxp:] = 0.5 * (Gb[mxp:, mxp:] + H(Gb[mxp:, mxp:]))
            Ga[mxp:, mxp:] = 0.5 * (Ga[mxp:, mxp:] + H(Ga[mxp:, mxp:]))
            return S, Gb, Ga

        S, Gb, Ga = build(Wa)
        fresh = Gb_c is None or (bound > refresh_tol if policy == 'bound' else (diag_dev > refresh_tol or since >= period))
        if fresh:
            Gb[:mxp, :mxp] = H(V) @ Bop(V)
            Ga[:mxp, :mxp] = H(V) @ A(V)
            bound = 0.0
            nrefresh += 1
            since = 0
        else:
            Gb[:mxp, :mxp] = Gb_c
            Ga[:mxp, :mxp] = Ga_c
        if log is not None:
            Gt = H(V) @ Bop(V); At = H(V) @ A(V)
            log.append((it, use_ortho, np.linalg.norm(Gb[:mxp, :mxp] - Gt), np.linalg.norm(Ga[:mxp, :mxp] - At) / anorm, bound, fresh))

        def rr(Gb, Ga, use_ortho):
            m = Gb.shape[0]
            if use_ortho:
                lam, Z = no._eigh_upper(Ga)
                return 1, Z[:, :k].copy(), no._cp_from_eigvecs(Z, k), lam[:k]
            DinvR, info, rcond = no._chol_transform(Gb)
            if info != 0 or rcond < 5e-3:
                return 2, None, None, None
            lam, Z = no._eigh_upper(H(DinvR) @ (Ga @ DinvR))
            return 0, DinvR @ Z[:, :k], DinvR @ no._cp_from_eigvecs(Z, k), lam[:k]

        uo, Cx, Cp, eig = rr(Gb, Ga, use_ortho)
        if uo == 2:
            use_ortho = 1
            Wa, _ = no.ortho_drop(Wa, V, eps, eps, B)
            S, Gb2, Ga2 = build(Wa)
            Gb2[:mxp, :mxp] = Gb[:mxp, :mxp]; Ga2[:mxp, :mxp] = Ga[:mxp, :mxp]
            Gb, Ga = Gb2, Ga2
            uo, Cx, Cp, eig = rr(Gb, Ga, 1)
        use_ortho = uo
        X = S @ Cx
        P = S @ Cp
        AX = A(X)
        W = no.get_residual(X, AX, eig, A, B)
        res = no.get_residual_norm(W, eig, nev, anorm, bnorm)
        converged = 0
        for i in range(nev):
            if res[i] > tol:
                break
            converged += 1
        # cache for the next pass: C = [Cx | Cp_act]
        C = np.concatenate([Cx, Cp[:, converged:]], axis=1)
        Gb_c = H(C) @ (Gb @ C) if not use_ortho else np.eye(C.shape[1], dtype=dtype)
        Ga_c = H(C) @ (Ga @ C)
        Gb_c = 0.5 * (Gb_c + H(Gb_c)); Ga_c = 0.5 * (Ga_c + H(Ga_c))
        # growth of the cache error: E' = Cxp^H E Cxp (+ fresh rounding amplified by the same congruence)
        g = np.linalg.norm(C[:mxp, :], 2) ** 2
        gall = np.linalg.norm(C, 2) ** 2
        epsm = np.finfo(no._real(dtype)).eps
        bound = g * bound + 8 * epsm * max(gall, 1.0)
        # drift monitor: true diagonal of the X block (free in the residual kernel: X, BX, AX are streamed anyway)
        dB = np.max(np.abs(np.einsum('ij,ij->j', X.conj(), Bop(X)).real - np.diag(Gb_c)[:k].real))
        dA = np.max(np.abs(np.einsum('ij,ij->j', X.conj(), AX).real - np.diag(Ga_c)[:k].real)) / anorm
        diag_dev = max(dB, dA)
        since += 1
        if converged == nev:
            break
        it += 1
    return dict(eig=np.asarray(eig), res=res, X=X, iter=it, converged=converged, nrefresh=nrefresh)


def report(name, r, ref, log):
    nev = len(ref)
    err = np.max(np.abs(r["eig"][:nev] - ref) / np.abs(ref))
    eb = max(l[2] for l in log); ea = max(l[3] for l in log)
    print(f"{name}: passes {r['iter']} conv {r['converged']} refreshes {r['nrefresh']} max rel eig err {err:.2e} "
          f"max true cache err B {eb:.2e} A {ea:.2e} max bound {max(l[4] for l in log):.2e}")


if __name__ == "__main__":
    # C1: unpreconditioned 2-D Laplacian (enters ortho mode, soft-locks)
    g = (60, 60); n = 3600; nev, k = 10, 20
    A = no.op_stencil(g)
    X0 = pr.initial_block(n, k, 7)
    for rt in (1e-12, 1e300):
        log = []
        r = lobpcg_cached(A, X0, nev, 1e-8, 3000, refresh_tol=rt, log=log)
        report(f"lap2d plain refresh_tol={rt:g}", r, pr.laplacian_eigs(g, nev), log)
    r0 = no.lobpcg(A, X0, nev, 1e-8, 3000)
    print("   oracle passes", r0["iter"])
    # generalized pencil + Chebyshev (amplification regime)
    g = (14, 14, 14); n = 14 ** 3; nev, k = 8, 16
    A = no.op_stencil(g); b = pr.mass_diagonal(n); B = no.op_diag(b)
    T = no.op_chebyshev(A, 8, 0.3, 12.0)
    X0 = pr.initial_block(n, k, 7)
    r0 = no.lobpcg(A, X0, nev, 1e-8, 3000, B=B, T=T)
    for rt in (1e-12, 1e300):
        log = []
        r = lobpcg_cached(A, X0, nev, 1e-8, 3000, B=B, T=T, refresh_tol=rt, log=log)
        report(f"pencil cheb refresh_tol={rt:g}", r, r0["eig"][:nev], log)
        if rt > 1:
            print("   growth of true cache error per pass:", " ".join(f"{l[2]:.1e}" for l in log))
    print("   oracle passes", r0["iter"])
    # Jacobi + potential
    g = (16, 16, 16); n = 16 ** 3; nev, k = 6, 12
    pot = pr.harmonic_potential(g, 0.3)
    A = no.op_stencil(g, potential=pot); T = no.op_diag(1.0 / (6.0 + pot))
    X0 = pr.initial_block(n, k, 7)
    r0 = no.lobpcg(A, X0, nev, 1e-8, 3000, T=T)
    for rt in (1e-12, 1e300):
        log = []
        r = lobpcg_cached(A, X0, nev, 1e-8, 3000, T=T, refresh_tol=rt, log=log)
        report(f"trap jacobi refresh_tol={rt:g}", r, r0["eig"][:nev], log)
    print("   oracle passes", r0["iter"])
