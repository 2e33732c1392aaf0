#!/bin/bash
# int8 tensor path as the default (auto for n >= 2^18): full GPU tests, bench with all legs, ncu captures of the default kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 --durations=6 > gpurun_out/pytest_u.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_u.log
tail -12 gpurun_out/pytest_u.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_u.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_u.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['kernels'], d['e2e'], d['time_to_solution'], {k:v['ms_per_step'] for k,v in d['windows'].items()})
r=d['roofline']; print({k:r[k] for k in ('achieved','peak','frac','traffic','ms_per_launch','f64_equivalent','share_of_step')})
PY
cap() {  # name regex cmd...
  local name=$1 rx=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o /tmp/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "cap $name rc=$?"
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$name.csv 2>/dev/null
}
rm -f gpurun_out/ncu_raw_*.csv
cap ozgram oz_gram_cluster_kernel python tools/kernel_bench.py gramcols 4096000 600 300 gram_i8=1
cap oznn oz_nn_kernel python tools/kernel_bench.py nn 4096000 900 300 gram_i8=1
