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
