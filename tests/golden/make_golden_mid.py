"""Generates tests/golden/reference_runs_mid.npz: outputs of the UNMODIFIED reference (oracle/_ref) on the mid-size,
multi-tile cases of tests/golden/mid_cases.py.  Run HERE (container with /root/reference and oracle/_ref built):

    make -C oracle && python tests/golden/make_golden_mid.py [case ...]

Takes a few minutes of CPU.  The GPU box has no /root/reference; tests there read only the .npz.
"""
import sys
import time
from pathlib import Path

import numpy as np
import scipy.sparse as sp  # noqa: F401  (import before oracle/_ref is dlopen'ed)

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from lobpcg_b200 import problems as pr  # noqa: E402
from oracle import ref_bindings as rb  # noqa: E402
import mid_cases  # noqa: E402

OUT = Path(__file__).resolve().parent / "reference_runs_mid.npz"


def run_case(tag, c, out):
    g = c["grid"]
    n = g[0] * g[1] * g[2]
    dt = c["dtype"]
    X0 = mid_cases.x0(c)
    t0 = time.perf_counter()
    if c["kind"] == "ilobpcg":
        A = rb.op_bdg(g, dt, c["shift"], c["d"])
        B = rb.op_diag(np.concatenate([np.ones(n), -np.ones(n)]), dt)
        r = rb.solve(A, X0, c["nev"], c["tol"], c["maxit"], B=B, indefinite=True)
        out[f"run_{tag}_sig"] = r["sig"]
    else:
        if c["csr"]:
            A = rb.op_csr(*pr.laplacian_csr(g, dtype=dt, potential=c["pot"]))
        else:
            A = rb.op_stencil(g, dt, potential=c["pot"])
        B = rb.op_diag(pr.mass_diagonal(n), dt) if c["mass"] else None
        T = rb.op_cheb(A, c["cheb"]["degree"], c["cheb"]["lo"], c["cheb"]["hi"]) if c["cheb"] else None
        r = rb.solve(A, X0, c["nev"], c["tol"], c["maxit"], B=B, T=T)
    dt_s = time.perf_counter() - t0
    nev = c["nev"]
    out[f"run_{tag}_eig"] = r["eig"]
    out[f"run_{tag}_res"] = r["res"]
    out[f"run_{tag}_meta"] = np.array([r["iter"], r["converged"], nev, c["k"]])
    # B-orthonormality of the reference's own eigenvectors, as a yardstick for the GPU test
    X = r["X"][:, :nev]
    if c["kind"] == "ilobpcg":
        bx = np.concatenate([np.ones(n), -np.ones(n)])[:, None] * X
    elif c["mass"]:
        bx = pr.mass_diagonal(n).astype(X.dtype)[:, None] * X
    else:
        bx = X
    gram = X.conj().T @ bx
    out[f"run_{tag}_ortho_err"] = np.array([np.linalg.norm(gram - np.diag(np.diag(gram).real.round()))])
    print(f"{tag}: iter {r['iter']} conv {r['converged']}/{nev} in {dt_s:.1f} s; eig[:4] {r['eig'][:4]} "
          f"max res {r['res'][:nev].max():.2e} ortho err {out[f'run_{tag}_ortho_err'][0]:.2e}", flush=True)


if __name__ == "__main__":
    if not rb.available():
        raise SystemExit("build oracle/_ref first: make -C oracle")
    allc = mid_cases.cases()
    want = sys.argv[1:] or list(allc)
    out = dict(np.load(OUT)) if OUT.exists() else {}
    for tag in want:
        run_case(tag, allc[tag], out)
        np.savez_compressed(OUT, **out)
    print("wrote", OUT, sum(v.nbytes for v in out.values()) // 1024, "KiB", flush=True)
    import os
    os._exit(0)
