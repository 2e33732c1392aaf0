#!/bin/bash
# 1-GPU: full GPU tests, bench (own arm + reference arm), kernel benches of the hot kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 --durations=10 > gpurun_out/pytest_h.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_h.log
tail -25 gpurun_out/pytest_h.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
echo "bench rc=$?"
tail -3 gpurun_out/bench_h.err
cat gpurun_out/bench_h.json | cut -c1-3000
(
python tools/kernel_bench.py gramcols 4096000 600 300
python tools/kernel_bench.py nn 4096000 900 512
python tools/kernel_bench.py nn 4096000 900 600
python tools/kernel_bench.py csr 128 128
python tools/kernel_bench.py gram 1024000 300 upper dtype=c128
python tools/kernel_bench.py gram 4096000 600 mb=200 dtype=f32
python tools/kernel_bench.py nn 4096000 600 400 dtype=f32
) > gpurun_out/kb_h.jsonl 2>&1
cut -c1-400 gpurun_out/kb_h.jsonl
