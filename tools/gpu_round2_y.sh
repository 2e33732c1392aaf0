#!/bin/bash
# per-rank shapes of an 8-GPU run of C5 (n_local = 512 000) and the auto threshold (n = 2^18) on one GPU
mkdir -p gpurun_out
for g in 80 64; do
python bench.py --grid $g --steps 10 --warmup 4 --no-cpu --no-e2e > gpurun_out/bench_y_$g.json 2> gpurun_out/bench_y_$g.err
echo "bench grid $g rc=$?"; tail -2 gpurun_out/bench_y_$g.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_y_$g.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, {k:(v.get('ms_per_step')) for k,v in d['kernels'].items()}, d['time_to_solution'] and {k:d['time_to_solution'][k] for k in ('seconds','passes','converged','max_rel_eig_err_vs_analytic')}, {k:v['ms_per_step'] for k,v in d['windows'].items()}, d['int8_tensor_path'])
PY
done
