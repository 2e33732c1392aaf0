#!/bin/bash
# 1-GPU: new tests only (general indefinite RR, full-size configs, status), launch list of one bench run
mkdir -p gpurun_out
python -m pytest tests/test_gpu_solver_mid.py tests/test_gpu_full_size.py tests/test_gpu_solver.py -m gpu -q --maxfail=25 --durations=12 > gpurun_out/pytest_c.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_c.log
tail -40 gpurun_out/pytest_c.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02_c5.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-tts --no-cpu --no-windows --no-fp64-probe > gpurun_out/ncu_c.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu_c.log | cut -c1-300
