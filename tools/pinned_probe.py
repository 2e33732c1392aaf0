"""What a PINNED caller buffer would give the reference-facing call: direct DMA rates (h2d / d2h) from pinned host memory and the
cost of pinning an existing pageable block with cudaHostRegister, against the staged pageable path of csrc/hostcopy.cu:
    python tools/pinned_probe.py [GB]"""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from lobpcg_b200 import api

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gb * 1e9 / 8)
ctx = api.Context(0)
L = api.lib()
d = torch.empty(n, dtype=torch.float64, device="cuda")
hp = torch.empty(n, dtype=torch.float64).pin_memory()
hp.uniform_()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(hp, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    hp.copy_(d, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
print(json.dumps(dict(probe="pinned_direct", gb=gb, h2d_gbs=n * 8 / (t1 - t0) / 1e9, d2h_gbs=n * 8 / (t2 - t1) / 1e9)), flush=True)
h = np.random.default_rng(1).random(n)          # pageable, touched
for rep in range(2):
    t0 = time.perf_counter(); L.lb2_memcpy_h2d(ctx.h, d.data_ptr(), h.ctypes.data, n * 8); t1 = time.perf_counter()
    L.lb2_memcpy_d2h(ctx.h, h.ctypes.data, d.data_ptr(), n * 8); t2 = time.perf_counter()
print(json.dumps(dict(probe="pageable_staged_hostcopy", gb=gb, h2d_gbs=n * 8 / (t1 - t0) / 1e9, d2h_gbs=n * 8 / (t2 - t1) / 1e9)), flush=True)
rt = torch.cuda.cudart()
t0 = time.perf_counter(); rc = rt.cudaHostRegister(h.ctypes.data, n * 8, 0); t1 = time.perf_counter()
print(json.dumps(dict(probe="cudaHostRegister", gb=gb, rc=int(rc), seconds=t1 - t0, gbs=n * 8 / (t1 - t0) / 1e9)), flush=True)
if int(rc) == 0:
    ht = torch.from_numpy(h)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); L.lb2_memcpy_h2d(ctx.h, d.data_ptr(), h.ctypes.data, n * 8); t1 = time.perf_counter()
    print(json.dumps(dict(probe="registered_block_through_hostcopy (still staged)", h2d_gbs=n * 8 / (t1 - t0) / 1e9)), flush=True)
    t0 = time.perf_counter(); rt.cudaHostUnregister(h.ctypes.data); t1 = time.perf_counter()
    print(json.dumps(dict(probe="cudaHostUnregister", seconds=t1 - t0)), flush=True)
