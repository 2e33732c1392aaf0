#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "csr or stencil" > gpurun_out/pytest_j.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_j.log
tail -5 gpurun_out/pytest_j.log
(
for g in 128 160; do
for pipe in 0 1 2; do
python tools/kernel_bench.py csr $g 128 csr_order=0 csr_pipe=$pipe
python tools/kernel_bench.py csr $g 128 csr_order=512 csr_pipe=$pipe
python tools/kernel_bench.py csr $g 128 csr_order=0 csr_pipe=$pipe spmm_cols=8
python tools/kernel_bench.py csr $g 128 csr_order=512 csr_pipe=$pipe spmm_cols=8
done
done
) > gpurun_out/kb_j.jsonl 2>&1
cut -c1-330 gpurun_out/kb_j.jsonl
