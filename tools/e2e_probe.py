import sys, time; sys.path.insert(0, ".")
import numpy as np
from lobpcg_b200 import api, problems as pr
g, nev = 160, 150; k = 300; n = g**3
A = api.stencil_op((g, g, g), np.float64)
st = api._setup(A, None, n, k, nev, np.float64, 1e-8, 12, None, None, False, 0)
X = st.X(); X[:, :] = 0.0
rng = np.random.default_rng(0)
for j in range(k): X[:, j] = rng.random(n) - 0.5
for rep in range(2):
    t0 = time.perf_counter(); api.lib().d_lobpcg(st.ptr); t = time.perf_counter() - t0
    print(f"rep {rep}: d_lobpcg 12 passes: {t:.3f} s  iter={st.st.iter}", flush=True)
