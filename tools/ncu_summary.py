"""Turns the per-kernel `ncu --set full` raw pages exported on the GPU box (gpurun_out/ncu_raw_<name>.csv, written by
tools/gpu_round2_*.sh with `ncu -i … --page raw --csv`) into profiles/ncu_full_r02.md and profiles/ncu_traffic_r02.json.

    python tools/ncu_summary.py gpurun_out > profiles/ncu_full_r02.md
"""
import csv
import json
import re
import sys
from pathlib import Path

KEEP = [
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "derived__lts__lts2xbar_bytes.sum.per_second",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.min.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")
TITLES = {
    "gramcols": ("gram_wl_kernel, column-block products [X P W]^H [W | AW] (K2/K3 of the cached pass)", "python tools/kernel_bench.py gramcols 4096000 600 300"),
    "tallnn": ("tall_nn_persist_kernel (K6 projection, 128x128 tiles)", "python tools/kernel_bench.py nn 4096000 900 512"),
    "csr": ("csr_kernel (K1, general CSR, 128^3 7-point matrix, 128 columns)", "python tools/kernel_bench.py csr 128 128"),
    "zmma": ("gram_zmma_kernel (K2 c64, Hermitian 300 x 300, n = 1.024 M)", "python tools/kernel_bench.py gram 1024000 300 upper dtype=c128"),
    "nntc5": ("nn_tc5_kernel (K6 f32 projection on tcgen05, 600 -> 400)", "python tools/kernel_bench.py nn 4096000 600 400 dtype=f32"),
    "gramtc5": ("gram_tc5_kernel (K3 f32 on tcgen05, 600 x 200 rectangular)", "python tools/kernel_bench.py gram 4096000 600 mb=200 dtype=f32"),
    "resid": ("residual_kernel (K7/K8)", "python tools/kernel_bench.py resid 4096000 300"),
    "ozgram": ("oz_gram_kernel (opt-in int8 tensor path: column-block Gram on tcgen05 kind::i8, lock-step cohorts)", "python tools/kernel_bench.py gramcols 4096000 600 300 gram_i8=1"),
    "oznn": ("oz_nn_kernel (opt-in int8 tensor path: projection 900 -> 300, S slices as MN-major operand)", "python tools/kernel_bench.py nn 4096000 900 300 gram_i8=1"),
    "ozsplit": ("oz_split_kernel (f64 block -> 7 int8 slices, tiled layout)", "python tools/kernel_bench.py gramcols 4096000 600 300 gram_i8=1"),
}


def main(d):
    d = Path(d)
    traffic = {}
    print("# ncu --set full summaries, round 2 (one launch per kernel, B200, `--clock-control none`)\n")
    print("Captured by `tools/gpu_round2_f.sh` / `tools/gpu_round2_p.sh` after the same command had run without ncu in the same call; raw pages exported on "
          "the box (`ncu -i … --page raw --csv`), summarised by `tools/ncu_summary.py`.  Durations are profiler-side (cold "
          "caches, serialised) — the bench numbers are CUDA-event timings.\n")
    for name, (title, cmd) in TITLES.items():
        f = d / f"ncu_raw_{name}.csv"
        if not f.exists():
            continue
        rows = list(csv.reader(open(f)))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        print(f"## {title}\n\n`{cmd}`\n\n`{m['Kernel Name'][0][:170]}`\n")
        print("| metric | value |\n|---|---|")
        for k in KEEP:
            if k in m and m[k][0] != "":
                print(f"| {k} | {m[k][0]} {m[k][1]} |")
        st = sorted(((float(v[0]), STALL.match(h).group(1)) for h, v in m.items() if STALL.match(h) and v[0] not in ("", "0")), reverse=True)
        print("| top stall reasons (warps per issue) | " + ", ".join(f"{n} {x:.2f}" for x, n in st[:6]) + " |\n")
        def gb(key):
            v, u = m.get(key, ("0", "byte"))
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
            return float(v or 0) * mult
        traffic[name] = {"kernel": m["Kernel Name"][0][:120], "dram_bytes_read": gb("dram__bytes_read.sum"),
                         "dram_bytes_write": gb("dram__bytes_write.sum"), "gpu_time_ms": float(m["gpu__time_duration.sum"][0]) *
                         {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(m["gpu__time_duration.sum"][1], 1)}
    tf = Path("profiles/ncu_traffic_r02.json")
    old = json.loads(tf.read_text()) if tf.exists() else {}
    old.update(traffic)                      # kernels that were not captured again keep their earlier record
    if "gramcols" in traffic:
        old["gram_wl_kernel_cols"] = traffic["gramcols"]
    tf.write_text(json.dumps(old, indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")
