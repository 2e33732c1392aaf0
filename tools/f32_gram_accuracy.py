import sys; sys.path.insert(0, ".")
import numpy as np
from lobpcg_b200 import api
ctx = api.Context(0)
rng = np.random.default_rng(0)
for n in (65536, 1048576, 4194304):
    A = np.asfortranarray(rng.standard_normal((n, 128)).astype(np.float32))
    dA = api.DeviceArray.from_numpy(ctx, A)
    ref = A.astype(np.float64).T @ A.astype(np.float64)
    for tc in (0, 1):
        ctx.set_option("gram_tc5", tc)
        G = api.gram(ctx, dA, dA, upper=True).numpy(ctx).astype(np.float64)
        d = (np.diag(G) - np.diag(ref)) / np.diag(ref)
        off = np.abs(G - ref).max() / np.abs(ref).max()
        print(f"n={n} tc5={tc} diag rel err mean {d.mean():+.3e} max|.| {np.abs(d).max():.3e}  max abs err / max entry {off:.3e}")

# projection (tall x small) in float: accumulation over kd columns
ctx.set_option("gram_tc5", -1)
for kd in (300, 900):
    n = 200000
    S = np.asfortranarray(np.abs(rng.standard_normal((n, kd))).astype(np.float32))      # all positive: worst case for a bias
    Cm = np.asfortranarray(np.abs(rng.standard_normal((kd, 64))).astype(np.float32))
    O = api.DeviceArray((n, 64), np.float32)
    api.tall_nn(ctx, api.DeviceArray.from_numpy(ctx, S), api.DeviceArray.from_numpy(ctx, Cm), O)
    ref = S.astype(np.float64) @ Cm.astype(np.float64)
    d = (O.numpy(ctx).astype(np.float64) - ref) / ref
    print(f"tall_nn f32 kd={kd}: rel err mean {d.mean():+.3e} max|.| {np.abs(d).max():.3e}")
# double Gram: is the DMMA accumulation unbiased?
n = 4194304
A = np.asfortranarray(np.abs(rng.standard_normal((n, 16))))
G = api.gram(ctx, api.DeviceArray.from_numpy(ctx, A), api.DeviceArray.from_numpy(ctx, A), upper=True).numpy(ctx)
ref = (A.astype(np.longdouble).T @ A.astype(np.longdouble))
d = ((G.astype(np.longdouble) - ref) / ref).astype(np.float64)
print(f"gram f64 n={n}: rel err mean {d.mean():+.3e} max|.| {np.abs(d).max():.3e}")
