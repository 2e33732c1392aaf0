"""Pageable host <-> device copy rate of csrc/hostcopy.cu against the number of memcpy threads: python tools/hostcopy_probe.py [GB]"""
import sys, time, json, os
sys.path.insert(0, ".")
import numpy as np
from lobpcg_b200 import api
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gb * 1e9 / 8)
ctx = api.Context(0)
h = np.random.default_rng(1).random(n)
h2 = np.zeros(n)
d = api.DeviceArray((n,), np.float64)
L = api.lib()
print(json.dumps(dict(cpus=os.cpu_count())))
for th in (4, 8, 12, 16, 24, 32):
    ctx.set_option("hostcopy_threads", th)
    for rep in range(2):
        t0 = time.perf_counter(); L.lb2_memcpy_h2d(ctx.h, d.ptr, h.ctypes.data, n * 8); t1 = time.perf_counter()
        L.lb2_memcpy_d2h(ctx.h, h2.ctypes.data, d.ptr, n * 8); t2 = time.perf_counter()
    print(json.dumps(dict(threads=th, h2d_gbs=n * 8 / (t1 - t0) / 1e9, d2h_gbs=n * 8 / (t2 - t1) / 1e9)), flush=True)
assert np.array_equal(h, h2)
