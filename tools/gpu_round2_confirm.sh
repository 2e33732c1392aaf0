#!/bin/bash
# confirmation run of the round's last build on a fresh box: full GPU tests, smoke, default bench
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 --durations=8 > gpurun_out/pytest_confirm.log 2>&1
echo "pytest rc=$? wall ${SECONDS}s" | tee -a gpurun_out/pytest_confirm.log
tail -14 gpurun_out/pytest_confirm.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
T0=$SECONDS
timeout 400 python bench.py > gpurun_out/bench_confirm.json 2> gpurun_out/bench_confirm.err
echo "bench rc=$? wall $((SECONDS-T0))s"; tail -3 gpurun_out/bench_confirm.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_confirm.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'])
print({k:v.get('ms_per_step') for k,v in d['kernels'].items()}, {k:v['ms_per_step'] for k,v in d['windows'].items()}, d['time_to_solution']['seconds'])
PY
