#!/bin/bash
# isolated timings of every hot kernel at its BASELINE-config shape on the round's last build, then the ncu launch list
# of the bench command (after the same command exited 0 without ncu)
mkdir -p gpurun_out
O=gpurun_out/kernel_bench_final_r02.jsonl; : > $O
SECONDS=0
kb() { timeout 60 python tools/kernel_bench.py "$@" >> $O 2>> gpurun_out/kernel_bench_final.err || echo "{\"failed\": \"$*\"}" >> $O; }
kb stencil 160 300
kb csr 128 128
kb resid 4096000 300
kb gramcols 4096000 600 300
kb gramcols 4096000 600 300 gram_i8=0
kb nn 4096000 900 300
kb nn 4096000 900 300 gram_i8=0
kb gram 4096000 900 upper
kb gram 4096000 600 upper dtype=f32
kb nn 4096000 600 200 dtype=f32
kb gram 1024000 300 upper dtype=c128
kb nn 1024000 300 100 dtype=c128
echo "kernel bench wall ${SECONDS}s"; cat $O | cut -c1-260
T0=$SECONDS
timeout 90 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-tts --no-windows > gpurun_out/bench_pre_ncu.json 2> gpurun_out/bench_pre_ncu.err
echo "bench rc=$? wall $((SECONDS-T0))s"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-tts --no-windows > gpurun_out/ncu_c.log 2>&1
echo "ncu rc=$? wall ${SECONDS}s"
