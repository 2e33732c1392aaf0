#!/bin/bash
mkdir -p gpurun_out
(
timeout 120 python tools/i8_debug.py 20000 150 150 upper
timeout 120 python tools/i8_debug.py 100003 300 100
) > gpurun_out/i8_debug.log 2>&1
grep -v "^gram_i8 [0-9]" gpurun_out/i8_debug.log | tail
(
timeout 300 python tools/kernel_bench.py gram 4096000 128 mb=128 gram_i8=2
timeout 300 python tools/kernel_bench.py gram 4096000 640 mb=256 gram_i8=2
timeout 300 python tools/kernel_bench.py gram 4096000 600 mb=300 gram_i8=2
timeout 300 python tools/kernel_bench.py gram 4096000 600 mb=300 gram_i8=2 oz_load_pct=60
timeout 300 python tools/kernel_bench.py gram 4096000 600 mb=300 gram_i8=2 oz_load_pct=130
timeout 300 python tools/kernel_bench.py gram 4096000 896 upper gram_i8=2
) > gpurun_out/kb_l.jsonl 2>&1
grep -v "^gram_i8" gpurun_out/kb_l.jsonl | cut -c1-300; grep "^gram_i8" gpurun_out/kb_l.jsonl | awk 'NR%7==0'
