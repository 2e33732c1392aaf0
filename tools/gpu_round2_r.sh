#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gram_cols or gram" > gpurun_out/pytest_r.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r.log
tail -5 gpurun_out/pytest_r.log
(
python tools/kernel_bench.py gramcols 4096000 600 300
python tools/kernel_bench.py gramcols 4096000 600 300 gram_merge=0
python tools/kernel_bench.py gramcols 4096000 600 300 gram_load_pct=60
python tools/kernel_bench.py gramcols 4096000 600 300 gram_load_pct=80
python tools/kernel_bench.py gramcols 4096000 450 225
python tools/kernel_bench.py gramcols 4096000 450 225 gram_merge=0
) > gpurun_out/kb_r.jsonl 2>&1
cut -c1-330 gpurun_out/kb_r.jsonl
