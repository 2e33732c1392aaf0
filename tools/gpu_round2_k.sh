#!/bin/bash
mkdir -p gpurun_out
(
for lp in 30 40 50 60 70 85; do python tools/kernel_bench.py gramcols 4096000 600 300 gram_load_pct=$lp; done
python tools/kernel_bench.py gramcols 4096000 600 300 gram_bk=16
python tools/kernel_bench.py gramcols 4096000 600 300 single
python tools/kernel_bench.py gramcols 4096000 400 200
python tools/kernel_bench.py gramcols 4096000 512 256
) > gpurun_out/kb_k.jsonl 2>&1
cut -c1-330 gpurun_out/kb_k.jsonl
