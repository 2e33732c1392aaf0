#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err
echo "bench rc=$? wall ${SECONDS}s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_last.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'])
PY
