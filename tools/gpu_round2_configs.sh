#!/bin/bash
# BASELINE configs C1..C4 at full size, unpreconditioned (as BASELINE.json names them), on the round's last build
mkdir -p gpurun_out
SECONDS=0
timeout 200 python tools/run_configs.py C1 C2 C2csr C3s C4 > gpurun_out/configs_full_size_r02.jsonl 2> gpurun_out/configs_r02.err
echo "rc=$? wall ${SECONDS}s"
timeout 200 python tools/run_configs.py C3d >> gpurun_out/configs_full_size_r02.jsonl 2>> gpurun_out/configs_r02.err
echo "rc=$? wall ${SECONDS}s"
cut -c1-420 gpurun_out/configs_full_size_r02.jsonl; tail -3 gpurun_out/configs_r02.err
