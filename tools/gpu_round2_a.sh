#!/bin/bash
# first GPU pass of round 2: tests, kernel timings of the column-block Gram, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 -x --deselect tests/test_gpu_reftests.py > gpurun_out/pytest_a.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_a.log
tail -30 gpurun_out/pytest_a.log
(
python tools/kernel_bench.py gramcols 4096000 600 300
python tools/kernel_bench.py gramcols 4096000 600 300 notri
python tools/kernel_bench.py gramcols 4096000 600 300 single
python tools/kernel_bench.py gramcols 4096000 300 300
python tools/kernel_bench.py gramcols 512000 600 300
python tools/kernel_bench.py gram 4096000 600 mb=300
python tools/kernel_bench.py gram 4096000 600 mb=300 gram_wl=1
python tools/kernel_bench.py gram 4096000 900 upper
python tools/kernel_bench.py gram 4096000 600 upper
python tools/kernel_bench.py nn 4096000 900 300
python tools/kernel_bench.py nn 4096000 900 600
) > gpurun_out/kb_a.jsonl 2>&1
cat gpurun_out/kb_a.jsonl
python bench.py --steps 6 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
echo "bench rc=$?"
tail -c 6000 gpurun_out/bench_a.json
tail -5 gpurun_out/bench_a.err
