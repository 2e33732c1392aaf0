#!/bin/bash
mkdir -p gpurun_out
python tools/kernel_bench.py gram 4096000 640 mb=256 gram_i8=1 > gpurun_out/plain_n.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gram_kernel -s 2 -c 1 -f -o /tmp/prof_oz python tools/kernel_bench.py gram 4096000 640 mb=256 gram_i8=1 > gpurun_out/ncu_oz.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_oz.ncu-rep --page raw --csv > gpurun_out/ncu_raw_oz.csv 2>/dev/null
ls -la gpurun_out/ncu_raw_oz.csv
