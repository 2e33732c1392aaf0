#!/bin/bash
# A/B of the staggered copy issue (nn_stagger) in the f64 tall_nn / work-list Gram kernels + their parity tests
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tall_nn or gram" > gpurun_out/pytest_i.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_i.log
tail -5 gpurun_out/pytest_i.log
(
for st in 1 0; do
python tools/kernel_bench.py gramcols 4096000 600 300 nn_stagger=$st
python tools/kernel_bench.py gram 4096000 900 upper nn_stagger=$st
python tools/kernel_bench.py gram 4096000 600 upper nn_stagger=$st
done
) > gpurun_out/kb_i.jsonl 2>&1
cut -c1-330 gpurun_out/kb_i.jsonl
