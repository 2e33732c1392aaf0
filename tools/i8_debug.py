"""Debug / accuracy probe of the int8 Ozaki Gram (gram_i8.cu): python tools/i8_debug.py N MA MB [upper]"""
import sys
sys.path.insert(0, ".")
import numpy as np
from lobpcg_b200 import api

n, ma, mb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
upper = len(sys.argv) > 4 and sys.argv[4] == "upper"
ctx = api.Context(0)
rng = np.random.default_rng(5)
A = rng.standard_normal((n, ma)) * np.exp(rng.uniform(-8, 8, ma))[None, :]
B = A if upper else rng.standard_normal((n, mb)) * np.exp(rng.uniform(-8, 8, mb))[None, :]
dA = api.DeviceArray.from_numpy(ctx, np.asfortranarray(A))
dB = dA if upper else api.DeviceArray.from_numpy(ctx, np.asfortranarray(B))
G = api.DeviceArray((ma, mb), np.float64)
L = api.lib()
ref = A.T @ B
# the reference itself has rounding errors of ~ sqrt(n) u |a||b|: use long double accumulation for small cases
if n * ma * mb < 2e9:
    ref = (A.astype(np.longdouble).T @ B.astype(np.longdouble)).astype(np.float64)
scale = np.sqrt(np.outer((A * A).sum(0), (B * B).sum(0)))
for i8 in (0, 1):
    ctx.set_option("gram_i8", i8)
    rc = L.lb2_d_gram(ctx.h, n, ma, mb, dA.ptr, dA.ld, dB.ptr, dB.ld, G.ptr, ma, int(upper))
    ctx.sync()
    g = G.numpy(ctx)
    err = np.abs(g - ref) / scale
    print(f"gram_i8={i8} rc={rc} max |err| / (|a||b|) = {err.max():.3e}  median {np.median(err):.3e}  nan {np.isnan(g).sum()}")
    if i8 and err.max() > 1e-12:
        bad = np.argwhere(err > 1e-12)
        print("bad entries:", len(bad), "first", bad[:5].tolist(), "rows range", bad[:, 0].min(), bad[:, 0].max(), "cols", bad[:, 1].min(), bad[:, 1].max())
        i, j = bad[0]
        print("got", g[i, j], "ref", ref[i, j], "ratio", g[i, j] / ref[i, j] if ref[i, j] else None)
