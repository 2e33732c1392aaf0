"""Where the time of one reference-facing d_lobpcg(alg) call goes (LB2_TIMING split on stderr), for a fresh random X0 and for
an advanced iterate as X0 (what bench.py's e2e leg uses):  python tools/e2e_probe2.py [grid] [nev] [passes]"""
import os, sys, time, json
sys.path.insert(0, ".")
os.environ["LB2_TIMING"] = "1"
import numpy as np
from lobpcg_b200 import api, problems as pr

g = int(sys.argv[1]) if len(sys.argv) > 1 else 160
nev = int(sys.argv[2]) if len(sys.argv) > 2 else 150
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 12
n, k = g ** 3, 2 * nev
A = api.stencil_op((g, g, g), np.float64)
api.lobpcg(api.stencil_op((24, 24, 24), np.float64), pr.initial_block(24 ** 3, 8, 7), 4, 1e-8, 5)   # library warm-up
ctx = api.Context(0)
for label, adv in (("fresh X0", 0), ("advanced iterate", 25)):
    s = api.Solver(ctx, A, n, k, nev, np.float64, 1e-8, 10 ** 6, device_seed=7)
    s.init()
    if adv:
        s.step(adv)
    X = s.finish()["X"]
    s.close()
    ctx.trim()
    st = api._setup(A, None, n, k, nev, np.float64, 1e-8, passes, None, None, False, 0)
    st.X()[:, :] = X
    del X
    for rep in range(2):
        st.st.iter = 0
        t0 = time.perf_counter()
        api.lib().d_lobpcg(st.ptr)
        t = time.perf_counter() - t0
        print(json.dumps(dict(x0=label, rep=rep, passes=int(st.st.iter), seconds=t, iters_per_s=int(st.st.iter) / t,
                              status=int(api.lib().lb2_last_status()))), flush=True)
    st.free()
