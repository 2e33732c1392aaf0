#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_solver.py tests/test_gpu_solver_mid.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 10 --warmup 4 --no-cpu --no-e2e > gpurun_out/bench_z2.json 2> gpurun_out/bench_z2.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_z2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_z2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, {k:v.get('ms_per_step') for k,v in d['kernels'].items()}, {k:v['ms_per_step'] for k,v in d['windows'].items()}, d['time_to_solution']['seconds'], d['time_to_solution']['max_rel_eig_err_vs_analytic'])
PY
