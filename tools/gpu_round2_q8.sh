#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_q8.json 2> gpurun_out/bench_q8.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_q8.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, {k:v.get('ms_per_step') for k,v in d['kernels'].items()}, d['e2e'] and {k:d['e2e'][k] for k in ('value','gpus_used','status','seconds')}, d['time_to_solution'] and {k:d['time_to_solution'][k] for k in ('seconds','passes','converged','max_rel_eig_err_vs_analytic')}, {k:v['ms_per_step'] for k,v in d['windows'].items()})
PY
