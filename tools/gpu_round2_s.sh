#!/bin/bash
mkdir -p gpurun_out
(
timeout 300 python tools/kernel_bench.py nn 4096000 900 300 gram_i8=1 oz_reuse=1
timeout 300 python tools/kernel_bench.py nn 4096000 900 600 gram_i8=1 oz_reuse=1
timeout 300 python tools/kernel_bench.py nn 4096000 900 576 gram_i8=1 oz_reuse=1
) > gpurun_out/kb_s.jsonl 2>&1
cut -c1-220 gpurun_out/kb_s.jsonl
