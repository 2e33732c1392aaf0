#!/bin/bash
# everything with the int8 tensor path ON (LB2_GRAM_I8=1 reaches every solver; kernel tests set their own options)
mkdir -p gpurun_out
LB2_GRAM_I8=1 python -m pytest tests -m gpu -q --maxfail=25 --durations=6 -k "not reftests" > gpurun_out/pytest_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_t.log
tail -14 gpurun_out/pytest_t.log
python bench.py --steps 10 --warmup 4 --no-cpu --no-e2e --no-tts > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_t.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_t.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['kernels'])
for k,v in d['windows'].items(): print(k, v['ms_per_step'], v['ms'])
PY
