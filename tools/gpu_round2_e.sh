#!/bin/bash
# 1-GPU: e2e timing split, windowed CSR (cp.async fill) vs plain, ncu --set full captures of the hot kernels
mkdir -p gpurun_out
python tools/e2e_probe2.py > gpurun_out/e2e_probe_e.log 2>&1
cat gpurun_out/e2e_probe_e.log | grep -v "^F-Norm" | tail -12
(
python tools/kernel_bench.py csr 128 128
python tools/kernel_bench.py csr 128 128 csr_window=0
python tools/kernel_bench.py csr 128 256
python tools/kernel_bench.py csr 128 256 csr_window=0
) > gpurun_out/kb_e.jsonl 2>&1
cat gpurun_out/kb_e.jsonl
cap() {  # name regex cmd...
  local name=$1 rx=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "cap $name rc=$?"
}
cap gramcols gram_wl_kernel python tools/kernel_bench.py gramcols 4096000 600 300
cap tallnn tall_nn_persist python tools/kernel_bench.py nn 4096000 900 512
cap csr csr_kernel python tools/kernel_bench.py csr 128 128 csr_window=0
cap csrwin csr_win_kernel python tools/kernel_bench.py csr 128 128
cap zmma gram_zmma_kernel python tools/kernel_bench.py gram 1024000 300 upper dtype=c128
cap nntc5 nn_tc5_kernel python tools/kernel_bench.py nn 4096000 600 400 dtype=f32
cap resid residual_kernel python tools/kernel_bench.py resid 4096000 300
ls -la gpurun_out/*.ncu-rep
