#!/bin/bash
mkdir -p gpurun_out
python tools/e2e_probe2.py 160 150 20 > gpurun_out/e2e_probe_w.log 2>&1
grep -v "^F-Norm" gpurun_out/e2e_probe_w.log | tail -24
