#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "i8 or int8" 2>&1 | tail -4
LB2_GRAM_I8=1 timeout 900 python -m pytest tests/test_gpu_solver_mid.py tests/test_gpu_solver.py tests/test_gpu_reftests.py -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 8 --warmup 4 --no-cpu --no-e2e --no-tts > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_z.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_z.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, {k:v.get('ms_per_step') for k,v in d['kernels'].items()})
for k,v in d['windows'].items(): print(k, v['ms_per_step'], v['ms'])
PY
