#!/bin/bash
# the driver's bench command on the round's last commit
mkdir -p gpurun_out
SECONDS=0
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_last2.json 2> gpurun_out/bench_last2.err
echo "bench rc=$? wall ${SECONDS}s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_last2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['cpu_baseline']['value'])
PY
