#!/bin/bash
mkdir -p gpurun_out
(
for lp in 130 160 200 250 300; do
timeout 300 python tools/kernel_bench.py gramcols 4096000 600 300 gram_i8=2 oz_load_pct=$lp
done
timeout 300 python tools/kernel_bench.py gramcols 4096000 450 225 gram_i8=2 oz_load_pct=160
timeout 300 python tools/kernel_bench.py gramcols 4096000 450 225 gram_i8=2 oz_load_pct=250
timeout 300 python tools/kernel_bench.py gramcols 4096000 450 225 gram_i8=2 oz_cluster=0
) > gpurun_out/kb_s.jsonl 2>&1
grep -v "^gram_i8" gpurun_out/kb_s.jsonl | cut -c1-200; grep "^gram_i8" gpurun_out/kb_s.jsonl | awk 'NR%7==0'
