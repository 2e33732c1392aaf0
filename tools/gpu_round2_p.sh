#!/bin/bash
# 1-GPU: full GPU tests, bench (own arm), launch list of the bench, ncu --set full captures of the hot kernels (raw pages exported on the box)
mkdir -p gpurun_out
rm -f gpurun_out/ncu_raw_*.csv
python -m pytest tests -m gpu -q --maxfail=25 --durations=8 > gpurun_out/pytest_p.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_p.log
tail -16 gpurun_out/pytest_p.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_p.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['kernels'], d['e2e'], d['time_to_solution']['seconds'], {k:v['ms_per_step'] for k,v in d['windows'].items()})
PY
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-tts --no-windows > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-tts --no-windows > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
cap() {  # name regex cmd...
  local name=$1 rx=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o /tmp/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "cap $name rc=$?"
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$name.csv 2>/dev/null
}
cap gramcols gram_wl_kernel python tools/kernel_bench.py gramcols 4096000 600 300
cap tallnn tall_nn_persist python tools/kernel_bench.py nn 4096000 900 512
cap csr csr_kernel python tools/kernel_bench.py csr 128 128
cap ozgram oz_gram_kernel python tools/kernel_bench.py gramcols 4096000 600 300 gram_i8=1
cap oznn oz_nn_kernel python tools/kernel_bench.py nn 4096000 900 300 gram_i8=1
cap ozsplit oz_split_kernel python tools/kernel_bench.py gramcols 4096000 600 300 gram_i8=1
du -sh gpurun_out
