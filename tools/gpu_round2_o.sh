#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "i8" > gpurun_out/pytest_o.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_o.log
tail -4 gpurun_out/pytest_o.log
(
timeout 300 python tools/kernel_bench.py nn 4096000 900 300 gram_i8=1
) > gpurun_out/kb_o.jsonl 2>&1
cut -c1-330 gpurun_out/kb_o.jsonl
