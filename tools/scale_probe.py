"""Phase profile of a few LOBPCG passes at a given size: python tools/scale_probe.py g nev steps [csr]"""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
from lobpcg_b200 import api, problems as pr

g = int(sys.argv[1]); nev = int(sys.argv[2]); steps = int(sys.argv[3]); kind = sys.argv[4] if len(sys.argv) > 4 else "stencil"
k = 2 * nev; n = g ** 3
ctx = api.Context()
if kind == "csr":
    rp, c, v = pr.laplacian_csr((g, g, g)); A = api.csr_op(rp, c, v)
else:
    A = api.stencil_op((g, g, g), np.float64)
T = None
s = api.Solver(ctx, A, n, k, nev, np.float64, 1e-8, 100000, device_seed=7)
t0 = time.time(); s.init(); ctx.sync(); t1 = time.time()
print(f"n={n} k={k} init {t1 - t0:.3f}s", s.stats(), flush=True)
s.reset_stats()
s.step(2); ctx.sync()
s.reset_stats()
t0 = time.time(); done = s.step(steps); ctx.sync(); t1 = time.time()
st = s.stats()
print(f"{done} passes in {t1 - t0:.3f}s = {(t1 - t0) / done * 1e3:.1f} ms/pass", s.progress())
tot = 0
for name, d in st.items():
    if d["ms"] <= 0: continue
    rate = d["work"] / d["ms"] / 1e9
    unit = "TFLOP/s" if name in ("gram", "tall_nn") else "TB/s"
    print(f"  {name:12s} {d['ms'] / done:9.2f} ms/pass  calls/pass {d['calls'] / done:6.1f}  {rate:8.3f} {unit}")
    tot += d["ms"]
print(f"  sum phases {tot / done:.2f} ms/pass")
