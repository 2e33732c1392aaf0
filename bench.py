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


def fp64_peak_probe(stream, device_index, seconds=1.0):
    """FP64 peak measured IN THIS RUN: cuBLAS DGEMM 8192^3 (through torch.matmul) back to back for ~`seconds`, CUDA events
    on the launching stream, SM clocks sampled during the probe.  A probe of the library peak, not a product path."""
    import torch
    n = 8192
    with torch.cuda.stream(stream):
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        c = torch.empty(n, n, dtype=torch.float64, device="cuda")
        for _ in range(2):
            torch.matmul(a, b, out=c)
        stream.synchronize()
        sampler = ClockSampler(device_index)
        sampler.start()
        time.sleep(0.25)
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 0
        e0.record(stream)
        while True:
            for _ in range(4):
                torch.matmul(a, b, out=c)
            iters += 4
            stream.synchronize()
            if time.time() - t0 >= seconds:
                break
        e1.record(stream)
        stream.synchronize()
        t1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1)
    del a, b, c
    return {"tflops": 2.0 * n ** 3 * iters / (ms * 1e-3) / 1e12, "iters": iters, "ms": ms, "clocks": clocks,
            "what": "cuBLAS DGEMM 8192^3 via torch.matmul(float64), back to back, measured inside this bench run"}


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


def bench_config(g, nev, k, world):
    """`config` of the JSON line: the SAME dict for both arms (the reference arm runs "on your arm's config"); what is
    specific to one run (solver state, grid really timed by the reference arm) goes into separate keys of the line."""
    n_local = g ** 3 // max(world, 1)
    return {"workload": workload_desc(g, nev, k), "parallelism": f"rows in {world} z-slab(s)",
            "l2": "inputs larger than L2 (each n x 3k slab is %.1f GB per GPU)" % (n_local * 3 * k * 8 / 1e9)}


# ------------------------------------------------------------------------------------------------ CPU legs
def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 0


def _ref_time_passes(rb, A, X0, nev, passes):
    """Seconds per pass of the UNMODIFIED reference: T(maxIter = passes) - T(maxIter = 0) over `passes` (the reference
    has no timers of its own; its set-up — allocation, ||A|| estimate, initial Rayleigh-Ritz — is in both runs)."""
    t0 = time.perf_counter(); rb.solve(A, X0, nev, 1e-8, 0); t_init = time.perf_counter() - t0
    t0 = time.perf_counter(); r = rb.solve(A, X0, nev, 1e-8, passes); t_run = time.perf_counter() - t0
    done = max(int(r["iter"]), 1)
    loop = t_run - t_init
    if loop < 0.05 * t_run:      # tiny samples: the difference of two short timings is noise
        loop = t_run
    return loop / done, done, t_init, t_run


def cpu_reference_rate(g_full, nev, k, steps, threads=None, g_sample=None):
    """Bounded sample for the `cpu_baseline` block of our own line: the unmodified reference on a g_s^3 grid with the
    same nev / sizeSub; iter/s scaled to the full row count and LABELLED as extrapolated (every O(n) term of a pass is
    linear in n; the O(k^3) small dense part stays unscaled in the measured time, which favours the CPU).  The
    reference ARM (--impl reference) runs the real configuration instead."""
    from oracle import ref_bindings as rb
    from lobpcg_b200 import problems as pr
    if not rb.available():
        return None
    threads = threads or os.cpu_count() or 1
    rb.set_threads(threads)
    g_s = g_sample or (64 if k >= 200 else 96)
    g_s = min(g_s, g_full)
    n_s = g_s ** 3
    A = rb.op_stencil((g_s, g_s, g_s), np.float64)
    X0 = pr.initial_block(n_s, k, 7)
    per_pass, passes, t_init, t_run = _ref_time_passes(rb, A, X0, nev, steps)
    scale = n_s / float(g_full ** 3)
    return dict(value=scale / per_pass, unit="iter/s", cores=threads, kind="reference",
                extrapolated=(g_s != g_full),
                sample=(f"unmodified reference (oracle/_ref, OpenBLAS {rb.blas_config().split()[1]}, {threads} threads) "
                        f"on a {g_s}^3 SAMPLE (n={n_s}) with the same nev={nev}, sizeSub={k}: {passes} passes in "
                        f"{t_run - t_init:.2f} s (set-up {t_init:.2f} s excluded); iter/s EXTRAPOLATED to n={g_full ** 3} "
                        f"by n_sample/n_full = {scale:.4f}" if g_s != g_full else
                        f"unmodified reference (oracle/_ref, OpenBLAS {rb.blas_config().split()[1]}, {threads} threads) "
                        f"on the full {g_full}^3 grid: {passes} passes in {t_run - t_init:.2f} s (set-up {t_init:.2f} s excluded)"),
                sample_seconds_per_pass=per_pass)


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation (unmodified sources compiled as oracle/_ref) on the SAME
    configuration — the real grid whenever host memory and the time budget allow, with as many passes of the requested
    K as fit the budget (>= 3).  Only when the full grid does not fit: the largest grid that does, flagged extrapolated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_bindings as rb
    from lobpcg_b200 import problems as pr
    g, nev = args.grid, args.nev
    k = 2 * nev
    if not rb.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_lobpcg.so missing (make -C oracle)"}))
        return
    t_wall = time.perf_counter()
    threads = os.cpu_count() or 1
    rb.set_threads(threads)
    # calibration on a small grid: seconds per pass and row, set-up seconds per row
    g_c = min(48, g)
    while 3 * k > g_c ** 3:
        g_c += 8
    Ac = rb.op_stencil((g_c, g_c, g_c), np.float64)
    per_pass_c, _, t_init_c, _ = _ref_time_passes(rb, Ac, pr.initial_block(g_c ** 3, k, 7), nev, 2)
    del Ac
    per_row, init_per_row = per_pass_c / g_c ** 3, t_init_c / g_c ** 3
    budget = float(args.ref_budget)

    def need_bytes(gg):      # reference footprint 12 n k scalars (lobpcg.h:597-608) + X0 + its copy in the harness + slack
        return int(14.5 * gg ** 3 * k * 8 * 1.1) + (2 << 30)

    def est_seconds(gg, passes):   # two runs (with and without the loop), each with the set-up
        return 2.0 * init_per_row * gg ** 3 + passes * per_row * gg ** 3

    avail = mem_available_bytes()
    g_run, why = g, None
    if need_bytes(g) > avail:
        why = f"host MemAvailable {avail / 2**30:.0f} GiB < {need_bytes(g) / 2**30:.0f} GiB needed for the full grid"
    elif est_seconds(g, 3) > budget:
        why = f"3 passes of the full grid estimated at {est_seconds(g, 3):.0f} s > budget {budget:.0f} s"
    if why:
        g_run = g
        while g_run > g_c and (need_bytes(g_run) > avail or est_seconds(g_run, 3) > budget):
            g_run -= 8
        g_run = max(g_run, g_c)
    n_run = g_run ** 3
    passes = int(max(3, min(max(args.steps, 1), (budget - 2.0 * init_per_row * n_run) / max(per_row * n_run, 1e-9))))
    A = rb.op_stencil((g_run, g_run, g_run), np.float64)
    X0 = pr.initial_block(n_run, k, 7)
    per_pass, done, t_init, t_run = _ref_time_passes(rb, A, X0, nev, passes)
    scale = n_run / float(g ** 3)
    value = scale / per_pass
    extrapolated = g_run != g
    sample = (f"unmodified reference (oracle/_ref, OpenBLAS {rb.blas_config().split()[1]}, {threads} threads) on "
              f"{'the full' if not extrapolated else 'a REDUCED'} {g_run}^3 grid (n={n_run}), nev={nev}, sizeSub={k}: {done} passes in "
              f"{t_run - t_init:.1f} s (set-up {t_init:.1f} s timed separately and excluded)"
              + (f"; iter/s EXTRAPOLATED to n={g ** 3} by n_run/n_full = {scale:.4f} because {why}" if extrapolated else ""))
    cfg = bench_config(g, nev, k, max(args.gpus, 1))
    if extrapolated:     # not the same configuration: say so where the two arms are compared
        cfg["workload"] += f" — reference arm timed on {g_run}^3 (n={n_run}) and extrapolated"
    line = {
        "impl": "reference", "metric": "lobpcg_iters_per_s", "value": value, "unit": "iter/s",
        "n_gpus": args.gpus, "steps": done, "steps_requested": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / value, "measured_ms_per_step": 1e3 * per_pass, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "extrapolated": extrapolated,
        "config": cfg,
        "reference_run": {"device": "host CPU", "grid_timed": g_run, "rows_timed": n_run, "passes_timed": done, "budget_s": budget,
                          "calibration": f"{g_c}^3: {per_pass_c:.3f} s/pass, set-up {t_init_c:.2f} s"},
        "cpu_baseline": {"value": value, "unit": "iter/s", "cores": threads, "kind": "reference", "sample": sample,
                         "extrapolated": extrapolated},
        "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_wall,
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

    # host-side barrier (gloo): waiting ranks must not park an NCCL kernel on their GPUs while rank 0 uses all devices
    host_grp = torch.distributed.new_group(backend="gloo") if world > 1 else None

    def host_barrier():
        if world > 1:
            torch.cuda.synchronize()
            torch.distributed.barrier(group=host_grp)

    s = api.Solver(ctx, A, n, k, nev, np.float64, 1e-8, 10 ** 6, device_seed=7)
    if part is not None:
        dist.attach(s, part)
    s.init()
    s.step(args.warmup)
    barrier()
    s.reset_stats()
    ctx.oz_stats()
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
    oz = ctx.oz_stats()                       # int8 tensor path: device time of split / MMA kernel / reduce inside the Gram phase
    int8_on = oz["calls"] > 0
    prog = s.progress()
    gram_cache_info = {"enabled": bool(s.info("gram_cache")), "refreshes_in_timed_region": None,
                       "monitor_max": s.info("gram_cache_monitor_max"), "arena_columns_per_k": s.info("arena_columns") / k,
                       "arena_gb": s.info("arena_bytes") / 1e9}
    if done != args.steps:
        raise SystemExit(f"solver stopped after {done} of {args.steps} passes (converged early?)")

    # ---- further timing windows on the same solver: ~50 % of the columns soft-locked, then the
    # sticky ortho branch (forced through a solver option); device time, max over ranks, like the main window
    windows = {}
    if not args.no_windows:
        def timed_window(name, nsteps, note):
            s.step(2)                      # settle into the new mode
            barrier()
            s.reset_stats()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                ea.record(stream)
                dn = s.step(nsteps)
                eb.record(stream)
            barrier()
            t = ea.elapsed_time(eb)
            if world > 1:
                tt = torch.tensor([t], device="cuda")
                torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
                t = float(tt.item())
            stw = s.stats()
            windows[name] = {"ms_per_step": t / max(dn, 1), "steps": dn, "note": note, "solver_state": s.progress(),
                             "ms": {kk: stw[kk]["ms"] / max(dn, 1) for kk in ("gram", "tall_nn", "spmm", "residual", "small_dense", "comm")}}
        nw_steps = max(3, min(args.steps, 6))
        # same state as the main window with the int8 tensor path switched off: the f64 DMMA kernels (gram_wl_kernel,
        # tall_nn_persist_kernel) that were the default before round 2's int8 path (lobpcg_b200/csrc/gram_i8.cu)
        if int8_on:
            ctx.set_option("gram_i8", 0)
            timed_window("dmma", nw_steps, "main-window pass on the FP64 tensor pipe (mma.sync DMMA work-list Gram + persistent projection; "
                         "context option gram_i8 = 0 / LB2_GRAM_I8=0)")
            ctx.set_option("gram_i8", -1)
        s.set_option("debug_min_conv", nev // 2)
        timed_window("softlocked", nw_steps, f"Cholesky branch with the leading {nev // 2} of {k} columns soft-locked "
                     "(forced: option debug_min_conv; P and W shrink to the active columns)")
        s.set_option("debug_min_conv", 0)
        s.set_option("force_ortho", 1)
        timed_window("ortho", nw_steps, "sticky ortho branch (forced: option force_ortho): ortho_drop(W | [X P]) + RR without the B-Gram")

    # per-kernel achieved rates over the timed region (phase timers = CUDA events on the solver stream)
    hbm, how_hbm, fp64, how64 = measured_peaks()
    probe = None
    if not args.no_fp64_probe:
        probe = fp64_peak_probe(stream, local_rank)
        fp64, how64 = probe["tflops"], probe["what"] + " (clocks: %s MHz median, reasons %s)" % (
            probe["clocks"].get("sm_mhz"), probe["clocks"].get("reasons"))
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
    tf = ROOT / "profiles" / "ncu_traffic_r02.json"
    if tf.exists() and g == 160 and nev == 150 and world == 1:   # the ncu capture is of this shape on one GPU
        try:
            t = json.loads(tf.read_text())["gram_wl_kernel_cols"]
            traffic = float(t["dram_bytes_read"] + t["dram_bytes_write"])
        except Exception:
            traffic = None
    gram_flops_launch = st["gram"]["work"] / max(st["gram"]["calls"], 1) / world
    if int8_on:
        # dominant kernel of the default path: the int8 Gram kernel.  Its arithmetic is 28 exact int8 slice products per f64
        # product (levels 6..12 of the 7 x 7 digit products), so a launch executes 28 x the algorithmic f64 flop count in int8
        # tensor operations.  Peak: MEASURED_PEAKS.json has no int8 entry; tcgen05.mma kind::i8 issues at twice the bf16 rate
        # on sm_100 (K = 32 against K = 16 per instruction at the same instruction time), so 2 x the measured bf16 rate.
        try:
            mp = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            i8_peak, i8_src = 2.0 * float(mp["bf16_tflops_sustained"]), "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (measured; no int8 entry in the file)"
        except Exception:
            i8_peak, i8_src = 2.0 * 1400.0, "2 x 1.4 PFLOP/s bf16 sustained, of fallback (B200_PROFILING.md)"
        mma_ms = oz["mma_ms"] / max(oz["calls"], 1)
        i8_tops = 28.0 * gram_flops_launch / (mma_ms * 1e-3) / 1e12 if mma_ms > 0 else 0.0
        try:
            t = json.loads(tf.read_text())["ozgram"] if (tf.exists() and g == 160 and nev == 150 and world == 1) else None
            traffic = float(t["dram_bytes_read"] + t["dram_bytes_write"]) if t else None
        except Exception:
            traffic = None
        roofline = {
            "kernel": "oz_gram_cluster_kernel (K2/K3 on tcgen05.mma kind::i8, SASS UTCIMMA: f64 column-block Gram [X P W]^H [W | A W] as 28 "
                      "exact int8 slice products of a 7-slice Ozaki split, TMA multicast operands, int32 accumulators in TMEM, int64 "
                      "partial sums; lobpcg_b200/csrc/gram_i8.cu, DESIGN.md 3b)",
            "bound": "tensor", "achieved": i8_tops, "peak": i8_peak, "unit": "TFLOP/s", "frac": i8_tops / i8_peak,
            "unit_note": "int8 tensor operations (exact integer multiply-adds x 2) per second of the MMA kernel alone",
            "traffic": traffic, "traffic_unit": "bytes per launch (dram read+write, ncu --set full, profiles/ncu_traffic_r02.json); "
                                                 "algorithmic operand bytes per launch = 7 n (m + n_w) = %.3g (slices of [X P W] and A W)" % (n_local * 7.0 * 4 * k),
            "peak_source": i8_src, "per_gpu": True,
            "algorithmic_flops_per_launch": 28.0 * gram_flops_launch,
            "algorithmic_flops_note": "28 int8 slice products x the needed f64 flops of the launch, 2 products x n x (2 m_xp n_w + n_w (n_w + 1)) = %.4g" % gram_flops_launch,
            "ms_per_launch": {"split": oz["split_ms"] / max(oz["calls"], 1), "mma_kernel": mma_ms, "reduce": oz["reduce_ms"] / max(oz["calls"], 1)},
            "f64_equivalent": {"tflops_gram_phase": gram_tf, "tflops_mma_kernel": gram_flops_launch / (mma_ms * 1e-3) / 1e12 if mma_ms > 0 else 0.0,
                               "fp64_dgemm_peak": fp64, "fp64_peak_source": how64,
                               "note": "the same products on the FP64 tensor pipe (window 'dmma'): gram_wl_kernel at 0.85 of this DGEMM peak"},
            "fp64_probe": probe,
            "share_of_step": oz["mma_ms"] / (ms if ms > 0 else 1.0),
        }
    else:
        roofline = {
            "kernel": "gram_wl_kernel (K2/K3, FP64 tensor pipe DMMA.8x8x4): per pass ONE launch for the W columns of both Grams, "
                      "[X P W]^H [W | A W]; the [X P] blocks come from the cached C^H G C (SURVEY 8f-2)",
            "bound": "tensor", "achieved": gram_tf, "peak": fp64, "unit": "TFLOP/s", "frac": gram_tf / fp64,
            "traffic": traffic, "traffic_unit": "bytes per launch (dram read+write, ncu --set full, profiles/ncu_traffic_r02.json); "
                                                 "algorithmic operand bytes per launch = n*(m + 2 n_w)*8 = %.3g" % (n_local * 5.0 * k * 8),
            "peak_source": how64, "per_gpu": True,
            "algorithmic_flops_per_launch": gram_flops_launch,
            "algorithmic_flops_note": "needed entries only: 2 products x n x (2 m_xp n_w + n_w (n_w + 1)); a recomputation of the "
                                      "cached blocks (every 64 passes) adds the Hermitian [X P] products",
            "fp64_probe": probe,
            "share_of_step": st["gram"]["ms"] / (ms if ms > 0 else 1.0),
        }

    # ---- e2e: the reference-facing call d_lobpcg(alg) with HOST buffers (X0 upload, result download inside) ----
    e2e = None
    if not args.no_e2e:
        passes = max(args.steps, 12)   # set-up (X0 upload, ||A||, initial RR, download) amortises over the passes
        st2 = None
        if world == 1:
            # the SAME X0 as the device-resident run above (splitmix64 seed 7, generated on the device and downloaded), so
            # the call walks through the same passes — the Cholesky-branch regime `value` is measured in
            st2 = api._setup(A, None, n, k, nev, np.float64, 1e-8, passes, None, None, False, 0)
            Xd = api.fill_uniform(ctx, n, k, np.float64, 7)
            api._ck(api.lib().lb2_memcpy_d2h(ctx.h, st2.st.S, Xd.ptr, n * k * 8), "d2h")
            Xd.free()
        s.close()
        ctx.sync()
        barrier()
        if rank == 0:
            # N GPUs: the SAME reference-facing call, d_lobpcg(alg) on host buffers, from this one process; the library
            # spreads it over the N devices itself (csrc/multigpu.cu: worker thread per device, rows of X0 / of the
            # eigenvectors move over N PCIe links).  The other ranks have released their arenas and wait.
            if world > 1:
                A_e2e = api.stencil_op((g, g, g), np.float64)
                st2 = api._setup(A_e2e, None, n, k, nev, np.float64, 1e-8, passes, None, None, False, 0)
                rng = np.random.default_rng(7)
                Xh = st2.X()
                for j0 in range(0, k, 25):
                    Xh[:, j0:j0 + 25] = rng.random((n, min(25, k - j0))) - 0.5
            api.lib().lb2_set_num_gpus(world)
            # warm-up of the reference-facing entry point on a small problem: one-time library initialisation of the default
            # contexts (cuSOLVER / cuBLAS handles, pinned staging rings, the in-process NCCL communicator) is not part of a step
            from lobpcg_b200 import problems as pr_w
            gw = 32 if world > 1 else 24
            api.lobpcg(api.stencil_op((gw, gw, gw), np.float64), pr_w.initial_block(gw ** 3, 8, 7), 4, 1e-8, 5)
            torch.cuda.synchronize()
            # ... and one untimed 1-pass call at the full size (the contract's warm-up applies to this leg too): the
            # library keeps the device allocation of a solve for the next one, a first call pays a 79 GB cudaMalloc
            x0_keep = np.array(st2.X()[:, :8], copy=True)       # (cheap check that X0 is restored below)
            st2.st.maxIter = 1
            api.lib().d_lobpcg(st2.ptr)
            st2.st.maxIter = passes
            st2.st.iter = 0
            st2.st.converged = 0
            if world == 1:
                Xd = api.fill_uniform(ctx, n, k, np.float64, 7)
                api._ck(api.lib().lb2_memcpy_d2h(ctx.h, st2.st.S, Xd.ptr, n * k * 8), "d2h")
                Xd.free()
                assert np.array_equal(x0_keep, st2.X()[:, :8])
            else:
                rng = np.random.default_rng(7)
                for j0 in range(0, k, 25):
                    Xh[:, j0:j0 + 25] = rng.random((n, min(25, k - j0))) - 0.5
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            api.lib().d_lobpcg(st2.ptr)
            t_e2e = time.perf_counter() - t0
            used = int(api.lib().lb2_last_num_gpus())
            api.lib().lb2_set_num_gpus(0)
            it = int(st2.st.iter)
            e2e = {"value": it / t_e2e, "unit": "iter/s", "passes": it, "seconds": t_e2e, "gpus_used": used,
                   "status": int(api.lib().lb2_last_status()),
                   "h2d_bytes_per_step": n * k * 8 / max(it, 1), "d2h_bytes_per_step": (n * k * 8) / max(it, 1) + (nev + k) * 8,
                   "warmup": "one untimed 1-pass d_lobpcg(alg) call at the same size (plus a 24^3 call for library start-up)",
                   "note": "whole d_lobpcg(alg) call on host buffers from ONE process" +
                           (f", spread over {used} GPUs inside the call (LB2_GPUS / lb2_set_num_gpus)" if world > 1 else "") +
                           f": X0 upload (pageable), ||A|| estimate, initial RR, {it} passes, eigenvector download; set-up "
                           "amortises over the pass count"}
            st2.free()
        host_barrier()
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
            "config": bench_config(g, nev, k, world), "solver_state": prog,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "time_to_solution": tts, "gpu_launches": int(launches),
            "clocks": clocks, "kernels": kernels, "hbm_peak_gbs": hbm, "hbm_peak_source": how_hbm,
            "windows": windows, "gram_cache": gram_cache_info,
            "dtype_note": ("f64 results; the tall products (Gram, projection) are evaluated EXACTLY on the int8 tensor cores from a 7-slice "
                           "split of the f64 operands (int32 / int64 accumulation, one f64 rounding per output), everything else in f64"
                           if int8_on else "f64 throughout (FP64 tensor pipe)"),
            "int8_tensor_path": {"active": bool(int8_on), "gram_ms_per_step": {kk: oz[kk] / max(done, 1) for kk in ("split_ms", "mma_ms", "reduce_ms")}},
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
    ap.add_argument("--ref-budget", type=float, default=300.0,
                    help="--impl reference: seconds for the two timed reference runs (set-up twice + the passes)")
    ap.add_argument("--no-windows", action="store_true", help="skip the ortho-mode and soft-locked timing windows")
    ap.add_argument("--no-fp64-probe", action="store_true", help="use the committed FP64 peak instead of probing it in this run")
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
