#!/bin/bash
# final build of round 2: full GPU tests, bench with every leg (own arm), reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 --durations=6 > gpurun_out/pytest_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -12 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['kernels'], d['e2e'], d['time_to_solution'], {k:v['ms_per_step'] for k,v in d['windows'].items()}, d['clocks'], d['cpu_baseline'])
r=d['roofline']; print({k:r[k] for k in ('achieved','peak','frac','traffic','ms_per_launch','share_of_step')})
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_final_ref.json
