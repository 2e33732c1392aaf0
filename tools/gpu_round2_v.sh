#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_v.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['kernels'], d['e2e']['value'], d['time_to_solution']['seconds'], {k:v['ms_per_step'] for k,v in d['windows'].items()})
r=d['roofline']; print({k:r[k] for k in ('achieved','peak','frac','traffic','ms_per_launch','share_of_step')})
PY
cap() {  # name regex cmd...
  local name=$1 rx=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o /tmp/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "cap $name rc=$?"
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$name.csv 2>/dev/null
}
cap ozgram oz_gram_cluster_kernel python tools/kernel_bench.py gramcols 4096000 600 300 gram_i8=1
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-tts --no-windows > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-tts --no-windows > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
