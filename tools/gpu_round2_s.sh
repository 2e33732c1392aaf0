#!/bin/bash
mkdir -p gpurun_out
(
for r in 3 4 5 6; do
timeout 300 python tools/kernel_bench.py gramcols 4096000 600 300 gram_i8=2 oz_ring=$r
done
for r in 5 6 7 8; do
timeout 300 python tools/kernel_bench.py nn 4096000 900 300 gram_i8=1 oz_reuse=1 oz_nn_ring=$r
done
) > gpurun_out/kb_s.jsonl 2>&1
grep -v "^gram_i8" gpurun_out/kb_s.jsonl | cut -c1-200; grep "^gram_i8" gpurun_out/kb_s.jsonl | awk 'NR%7==0'
