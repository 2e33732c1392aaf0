"""Isolated kernel timing (CUDA events on the launching stream, L2-exceeding inputs):
   python tools/kernel_bench.py stencil G NC | csr G NC | gram N M [upper] | nn N KD NB | resid N NC"""
import sys, json
sys.path.insert(0, ".")
import numpy as np
import torch
from lobpcg_b200 import api, problems as pr

mode = sys.argv[1]
DT = {"f32": np.float32, "f64": np.float64, "c64": np.complex64, "c128": np.complex128}[next((a.split("=")[1] for a in sys.argv if a.startswith("dtype=")), "f64")]
PFX = api.PREFIX[np.dtype(DT)]
CMUL = 4.0 if np.dtype(DT).kind == "c" else 1.0
stream = torch.cuda.Stream()
ctx = api.Context(0, stream.cuda_stream)
MB = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("mb=")), None)
PAD = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("pad=")), 0)
for kv in [a for a in sys.argv if "=" in a and not a.startswith(("dtype=", "mb=", "pad="))]:
    k, v = kv.split("="); ctx.set_option(k, int(v))
args = [a for a in sys.argv[2:] if "=" not in a]

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    ctx.sync()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); ctx.sync()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))

if mode in ("stencil", "csr"):
    g, nc = int(args[0]), int(args[1]); n = g ** 3
    X = api.fill_uniform(ctx, n, nc, np.float64, 1); Y = api.DeviceArray((n, nc), np.float64)
    if mode == "csr":
        import os
        os.environ["LB2_CSR_NO_STENCIL_DETECT"] = "1"     # the general CSR kernels, not the recognised stencil
        rp, c, v = pr.laplacian_csr((g, g, g)); op = api.csr_op(rp, c, v); extra = len(v) * 12 + 8 * (n + 1)
    else:
        op = api.stencil_op((g, g, g), np.float64); extra = 0
    med, mn = timeit(lambda: op.apply(ctx, X, Y))
    b = 2.0 * n * nc * 8 + extra
    print(json.dumps(dict(kernel=mode, g=g, nc=nc, ms=med, ms_min=mn, gbs=b / med / 1e6, frac_hbm=b / med / 1e6 / 6536.7)))
elif mode == "gram":
    n, m = int(args[0]), int(args[1]); upper = len(args) > 2 and args[2] == "upper"
    mb = MB or m
    distinct = "distinct" in args      # Hermitian product of two different blocks (S^H AS): upper, but two panels per tile
    A = api.fill_uniform(ctx, n, m, DT, 1, ld=n + PAD); B = A if (upper and not distinct) else api.fill_uniform(ctx, n, mb, DT, 2, ld=n + PAD)
    L = api.lib(); G = api.DeviceArray((m, mb), DT)
    fn = getattr(L, f"lb2_{PFX}_gram")
    med, mn = timeit(lambda: fn(ctx.h, n, m, mb, A.ptr, A.ld, B.ptr, B.ld, G.ptr, m, int(upper)))
    fl = CMUL * (n * m * (m + 1.0) if upper else 2.0 * n * m * mb)
    print(json.dumps(dict(kernel="gram", dtype=str(np.dtype(DT)), n=n, m=m, mb=mb, upper=upper, ms=med, ms_min=mn, pad=PAD, tflops=fl / med / 1e9, frac=fl / med / 1e9 / 35.76, gbs=8.0 * n * (m + (0 if upper else mb)) / med / 1e6)))
elif mode == "gramcols":   # [X P W]^H [W | AW]: gramcols N MXP NW [single] [notri]
    n, mxp, nw = int(args[0]), int(args[1]), int(args[2])
    m = mxp + nw
    S = api.fill_uniform(ctx, n, m, DT, 1); AW = api.fill_uniform(ctx, n, nw, DT, 2)
    W = S.rows(0, n); W = api.DeviceArray.__new__(api.DeviceArray); W.shape, W.dtype, W.ld, W.ptr, W.nbytes, W._owner = (n, nw), S.dtype, S.ld, S.ptr + mxp * S.ld * S.dtype.itemsize, 0, S
    G0 = api.DeviceArray((m, nw), DT); G1 = api.DeviceArray((m, nw), DT)
    L = api.lib(); fn = getattr(L, f"lb2_{PFX}_gram_cols")
    single = "single" in args; tri = -1 if "notri" in args else mxp
    med, mn = timeit(lambda: fn(ctx.h, n, m, nw, S.ptr, S.ld, W.ptr, W.ld, G0.ptr, m, None if single else AW.ptr, AW.ld, None if single else G1.ptr, m, tri))
    nprod = 1 if single else 2
    fl = CMUL * nprod * n * (2.0 * mxp * nw + nw * (nw + 1.0))
    print(json.dumps(dict(kernel="gram_cols", dtype=str(np.dtype(DT)), n=n, mxp=mxp, nw=nw, nprod=nprod, tri=tri, ms=med, ms_min=mn, tflops_needed=fl / med / 1e9, frac=fl / med / 1e9 / 35.76, tflops_rect=CMUL * nprod * 2.0 * n * m * nw / med / 1e9)))
elif mode == "nn":
    n, kd, nb = int(args[0]), int(args[1]), int(args[2])
    S = api.fill_uniform(ctx, n, kd, DT, 1, ld=n + PAD); Cm = api.fill_uniform(ctx, kd, nb, DT, 2); O = api.DeviceArray((n, nb), DT, ld=n + PAD)
    med, mn = timeit(lambda: api.tall_nn(ctx, S, Cm, O))
    fl = CMUL * 2.0 * n * kd * nb
    print(json.dumps(dict(kernel="tall_nn", dtype=str(np.dtype(DT)), n=n, kd=kd, nb=nb, ms=med, ms_min=mn, tflops=fl / med / 1e9, frac=fl / med / 1e9 / 35.76)))
elif mode == "resid":
    n, nc = int(args[0]), int(args[1])
    AX = api.fill_uniform(ctx, n, nc, np.float64, 1); BX = api.fill_uniform(ctx, n, nc, np.float64, 2)
    lam = api.DeviceArray.from_numpy(ctx, np.ones(nc)); W = api.DeviceArray((n, nc), np.float64); ss = api.DeviceArray((nc,), np.float64)
    L = api.lib()
    med, mn = timeit(lambda: L.lb2_d_residual(ctx.h, n, nc, AX.ptr, n, BX.ptr, n, lam.ptr, W.ptr, n, ss.ptr))
    b = 3.0 * n * nc * 8
    print(json.dumps(dict(kernel="residual", n=n, nc=nc, ms=med, gbs=b / med / 1e6, frac_hbm=b / med / 1e6 / 6536.7)))
