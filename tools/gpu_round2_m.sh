#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "i8" > gpurun_out/pytest_m.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_m.log
tail -5 gpurun_out/pytest_m.log
(
timeout 300 python tools/kernel_bench.py gramcols 4096000 600 300 gram_i8=2 oz_cluster=0
timeout 300 python tools/kernel_bench.py gramcols 4096000 600 300 gram_i8=2 oz_cluster=0 oz_lockstep=0 oz_load_pct=160
timeout 300 python tools/kernel_bench.py gramcols 4096000 512 256 gram_i8=2 oz_cluster=0
timeout 300 python tools/kernel_bench.py gram 4096000 640 mb=256 gram_i8=2
timeout 300 python tools/kernel_bench.py gram 4096000 896 upper gram_i8=2
) > gpurun_out/kb_m.jsonl 2>&1
grep -v "^gram_i8" gpurun_out/kb_m.jsonl | cut -c1-330; grep "^gram_i8" gpurun_out/kb_m.jsonl | awk 'NR%7==0'
