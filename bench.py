#!/usr/bin/env python
"""bench.py — LOBPCG hot-path benchmark on B200 (contract: see README / DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid G] [--nev NEV]

A "step" is one pass of the LOBPCG loop (reference src/core/lobpcg_impl.inc:130-245): RR on [X P W],
projection, A*X, residual + norms.  Workload = BASELINE.json config C5 (3-D 7-point Dirichlet Laplacian
160^3, matrix-free, nev=150, sizeSub=300, double).  N>1 is launched by torchrun, one rank per GPU, rows
partitioned in z-slabs (lobpcg_b200/dist.py).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FP64_PEAK_FILE = ROOT / "profiles" / "fp64_peak_probe_r01.jsonl"


def measured_peaks():
    hbm, how_hbm = 6650.0, "fallback"
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            hbm, how_hbm = float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    # FP64 is not in MEASURED_PEAKS.json (bf16 only): cuBLAS DGEMM 8192^3 measured on this pool by
    # tools/fp64_peak_probe.cu, committed under profiles/
    fp64, how64 = 35.76, "cuBLAS DGEMM 8192^3 measured by tools/fp64_peak_probe.cu (profiles/fp64_peak_probe_r01.jsonl)"
    if FP64_PEAK_FILE.exists():
        for line in FP64_PEAK_FILE.read_text().splitlines():
            try:
                d = json.loads(line)
            except Exception:
                continue
            if d.get("probe") == "cublas_dgemm_8192_sustained":
                fp64 = float(d["tflops"])
    return hbm, how_hbm, fp64, how64


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 - 0.05 <= t <= t1 + 0.05):
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_desc(g, nev, k):
    return (f"C5: 3-D 7-point Dirichlet Laplacian {g}^3 (n={g ** 3}), matrix-free stencil, nev={nev}, sizeSub={k}, "
            f"double, tol=1e-8, B=T=NULL, X0 splitmix64 seed 7")


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_reference_rate(g_full, nev, k, steps, threads=None):
    """Times the UNMODIFIED reference (oracle/_ref) on a bounded sample: same operator family and block
    sizes on a g_s^3 grid; iter/s is scaled to the full row count (every O(n) term of a pass is linear in n;
    the O(k^3) small dense part is kept unscaled in the measured time, which favours the CPU)."""
    from oracle import ref_bindings as rb
    from lobpcg_b200 import problems as pr
    if not rb.available():
        return None
    threads = threads or os.cpu_count() or 1
    rb.set_threads(threads)
    g_s = 64 if k >= 200 else 96
    g_s = min(g_s, g_full)
    n_s = g_s ** 3
    A = rb.op_stencil((g_s, g_s, g_s), np.float64)
    X0 = pr.initial_block(n_s, k, 7)
    t0 = time.perf_counter(); rb.solve(A, X0, nev, 1e-8, 0); t_init = time.perf_counter() - t0
    t0 = time.perf_counter(); r = rb.solve(A, X0, nev, 1e-8, steps); t_run = time.perf_counter() - t0
    passes = max(int(r["iter"]), 1)
    # tiny samples: the difference of two short timings is noise; fall back to the whole call
    per_pass = (t_run - t_init if t_run - t_init > 0.05 * t_run else t_run) / passes
    scale = n_s / float(g_full ** 3)
    return dict(value=scale / per_pass, unit="iter/s", cores=threads, kind="reference",
                sample=(f"unmodified reference (oracle/_ref, OpenBLAS {rb.blas_config().split()[1]}, {threads} threads) "
                        f"on {g_s}^3 rows with the same nev={nev}, sizeSub={k}: {passes} passes in {t_run - t_init:.2f} s "
                        f"(init {t_init:.2f} s excluded), iter/s scaled by n_sample/n_full = {scale:.4f}"),
                sample_seconds_per_pass=per_pass)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g, nev = args.grid, args.nev
    k = 2 * nev
    t0 = time.perf_counter()
    cb = cpu_reference_rate(g, nev, k, max(args.steps, 1))
    if cb is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_lobpcg.so missing (make -C oracle)"}))
        return
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": "lobpcg_iters_per_s", "value": cb["value"], "unit": "iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_desc(g, nev, k), "device": "host CPU"},
        "cpu_baseline": {k2: cb[k2] for k2 in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    from lobpcg_b200 import api, dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local_rank)
    stream = torch.cuda.Stream()
    ctx = api.Context(local_rank, stream.cuda_stream)
    grp = None
    if world > 1:
        grp = dist.init_process_group(ctx, backend="nccl")

    g, nev = args.grid, args.nev
    k = 2 * nev
    n = g ** 3
    if world > 1:
        part = dist.SlabPartition(g, g, g, world, rank)
        A = dist.partitioned_stencil(ctx, part, np.float64, k)
        n_local = part.n_local
    else:
        part = None
        A = api.stencil_op((g, g, g), np.float64)
        n_local = n

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    s = api.Solver(ctx, A, n, k, nev, np.float64, 1e-8, 10 ** 6, device_seed=7)
    if part is not None:
        dist.attach(s, part)
    s.init()
    s.step(args.warmup)
    barrier()
    s.reset_stats()
    l0 = ctx.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    with torch.cuda.stream(stream):
        e0.record(stream)
        done = s.step(args.steps)
        e1.record(stream)
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(tms, op=torch.distributed.ReduceOp.MAX)
        ms = float(tms.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    launches = ctx.launches - l0
    st = s.stats()
    prog = s.progress()
    if done != args.steps:
        raise SystemExit(f"solver stopped after {done} of {args.steps} passes (converged early?)")

    # per-kernel achieved rates over the timed region (phase timers = CUDA events on the solver stream)
    hbm, how_hbm, fp64, how64 = measured_peaks()
    def rate(name, unit_div):
        d = st[name]
        return (d["work"] / (d["ms"] * 1e-3) / unit_div) if d["ms"] > 0 else 0.0
    if world > 1:  # algorithmic work is per rank; aggregate over ranks
        for name in st:
            w = torch.tensor([st[name]["work"], st[name]["ms"]], device="cuda", dtype=torch.float64)
            wsum = w.clone(); torch.distributed.all_reduce(wsum, op=torch.distributed.ReduceOp.SUM)
            wmax = w.clone(); torch.distributed.all_reduce(wmax, op=torch.distributed.ReduceOp.MAX)
            st[name]["work"], st[name]["ms"] = float(wsum[0].item()), float(wmax[1].item())
    gram_tf = rate("gram", 1e12) / world
    kernels = {
        "gram": {"tflops": rate("gram", 1e12), "ms_per_step": st["gram"]["ms"] / done, "launches_per_step": st["gram"]["calls"] / done},
        "tall_nn": {"tflops": rate("tall_nn", 1e12), "ms_per_step": st["tall_nn"]["ms"] / done},
        "spmm": {"gbs": rate("spmm", 1e9), "ms_per_step": st["spmm"]["ms"] / done, "frac_hbm": rate("spmm", 1e9) / world / hbm},
        "residual": {"gbs": rate("residual", 1e9), "ms_per_step": st["residual"]["ms"] / done},
        "small_dense": {"ms_per_step": st["small_dense"]["ms"] / done},
        "comm": {"ms_per_step": st["comm"]["ms"] / done},
    }
    traffic = None
    tf = ROOT / "profiles" / "ncu_traffic_r01.json"
    if tf.exists() and g == 160 and nev == 150 and world == 1:   # the ncu capture is of this shape on one GPU
        t = json.loads(tf.read_text())["gram_wl_kernel"]
        traffic = float(t["dram_bytes_read"] + t["dram_bytes_write"])
    roofline = {
        "kernel": "gram_wl_kernel + gram_dmma_kernel strip (K2/K3: S^H S and S^H A S, FP64 tensor pipe DMMA.8x8x4)",
        "bound": "tensor", "achieved": gram_tf, "peak": fp64, "unit": "TFLOP/s", "frac": gram_tf / fp64,
        "traffic": traffic, "traffic_unit": "bytes per launch (dram read+write, ncu, profiles/ncu_traffic_r01.json); "
                                             "algorithmic operand bytes per launch = n*m*8 = %.3g" % (n_local * 3.0 * k * 8),
        "peak_source": how64, "per_gpu": True,
        "algorithmic_flops_per_launch": st["gram"]["work"] / max(st["gram"]["calls"], 1) / world,
        "share_of_step": st["gram"]["ms"] / (ms if ms > 0 else 1.0),
    }

    # ---- e2e: the reference-facing call d_lobpcg(alg) with HOST buffers (X0 upload, result download inside) ----
    e2e = None
    if world == 1 and not args.no_e2e:
        X_dev = s.finish()["X"]            # any non-zero host block will do as X0; reuse the current iterate
        passes = max(args.steps, 12)   # set-up (X0 upload, ||A||, initial RR, download) amortises over the passes
        st2 = api._setup(A, None, n, k, nev, np.float64, 1e-8, passes, None, None, False, 0)
        st2.X()[:, :] = X_dev
        s.close()
        ctx.sync()
        # warm-up of the reference-facing entry point on a small problem: one-time library initialisation of the default
        # context (cuSOLVER / cuBLAS handles, pinned staging ring) is not part of a step
        from lobpcg_b200 import problems as pr_w
        api.lobpcg(api.stencil_op((24, 24, 24), np.float64), pr_w.initial_block(24 ** 3, 8, 7), 4, 1e-8, 5)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        api.lib().d_lobpcg(st2.ptr)
        t_e2e = time.perf_counter() - t0
        it = int(st2.st.iter)
        e2e = {"value": it / t_e2e, "unit": "iter/s", "passes": it, "seconds": t_e2e,
               "h2d_bytes_per_step": n * k * 8 / max(it, 1), "d2h_bytes_per_step": (n * k * 8) / max(it, 1) + (nev + k) * 8,
               "note": "whole d_lobpcg(alg) call on host buffers: X0 upload (pageable), ||A|| estimate, initial RR, "
                       f"{it} passes, eigenvector download; set-up amortises over the pass count"}
        st2.free()
    else:
        s.close()

    # ---- time to solution of the same config (BASELINE metric "time-to-solution"): full solve to tol 1e-8 with the
    # built-in polynomial preconditioner behind alg->T; the unpreconditioned solve needs 793 passes
    # (profiles/full_solve_c5_final_r01.json) and does not fit a bench run
    tts = None
    if not args.no_tts:
        from lobpcg_b200 import problems as pr
        if world > 1:
            A2 = dist.partitioned_stencil(ctx, part, np.float64, k)
        else:
            A2 = api.stencil_op((g, g, g), np.float64)
        T2 = api.chebyshev_op(A2, args.cheb_degree, args.cheb_lo, 0.0, mixed=not args.cheb_full_precision)
        s2 = api.Solver(ctx, A2, n, k, nev, np.float64, 1e-8, 2000, T=T2, device_seed=7)
        if part is not None:
            dist.attach(s2, part)
        barrier()
        t0 = time.perf_counter()
        s2.init()
        while s2.step(10) == 10:
            pass
        barrier()
        t_tts = time.perf_counter() - t0
        p2 = s2.progress()
        eigs2, res2 = s2.results()
        an = pr.laplacian_eigs((g, g, g), nev)
        tts = {"seconds": t_tts, "passes": int(p2["iter"]) + 1, "converged": int(p2["converged"]), "nev": nev, "tol": 1e-8,
               "preconditioner": f"lb2_op_chebyshev{'' if args.cheb_full_precision else '_mixed'}(A, degree={args.cheb_degree}, "
                                 f"lo={args.cheb_lo}, hi=Gershgorin)",
               "max_rel_eig_err_vs_analytic": float(np.max(np.abs(eigs2[:nev] - an) / an)),
               "max_resnorm": float(res2[:nev].max()),
               "unpreconditioned_reference_point": "793 passes, 296 s on 1 GPU (profiles/full_solve_c5_final_r01.json)"}
        s2.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference_rate(g, nev, k, 2)
        if cpu:
            cpu = {k2: cpu[k2] for k2 in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": "lobpcg_iters_per_s", "value": done / (ms * 1e-3), "unit": "iter/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / done, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_desc(g, nev, k), "parallelism": f"rows in {world} z-slab(s)",
                       "l2": "inputs larger than L2 (each n x 3k slab is %.1f GB per GPU)" % (n_local * 3 * k * 8 / 1e9),
                       "solver_state": prog},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "time_to_solution": tts, "gpu_launches": int(launches),
            "clocks": clocks, "kernels": kernels, "hbm_peak_gbs": hbm, "hbm_peak_source": how_hbm,
        }
        print(json.dumps(line))
    if world > 1:
        dist.shutdown(ctx)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=160)
    ap.add_argument("--nev", type=int, default=150)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-tts", action="store_true", help="skip the preconditioned full solve (time_to_solution)")
    ap.add_argument("--cheb-degree", type=int, default=40)
    ap.add_argument("--cheb-lo", type=float, default=0.02)
    ap.add_argument("--cheb-full-precision", action="store_true", help="evaluate the preconditioner in double, not float")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
