#!/bin/bash
# 1-GPU: full GPU tests, e2e split, bench, ncu --set full captures exported to CSV on the box (the .ncu-rep files stay there)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python -m pytest tests -m gpu -q --maxfail=25 --deselect tests/test_gpu_reftests.py --durations=8 > gpurun_out/pytest_f.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_f.log
tail -22 gpurun_out/pytest_f.log
python tools/e2e_probe2.py 160 150 20 > gpurun_out/e2e_probe_f.log 2>&1
grep -v "^F-Norm" gpurun_out/e2e_probe_f.log | tail -10
python bench.py --steps 6 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_f.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['kernels'], d['e2e'], d['time_to_solution']['seconds'], d['gram_cache'], {k:v['ms_per_step'] for k,v in d['windows'].items()})
PY
cap() {  # name regex cmd...
  local name=$1 rx=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o /tmp/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "cap $name rc=$?"
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$name.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu_src_$name.csv.gz
}
cap gramcols gram_wl_kernel python tools/kernel_bench.py gramcols 4096000 600 300
cap tallnn tall_nn_persist python tools/kernel_bench.py nn 4096000 900 512
cap csr csr_kernel python tools/kernel_bench.py csr 128 128
cap zmma gram_zmma_kernel python tools/kernel_bench.py gram 1024000 300 upper dtype=c128
cap nntc5 nn_tc5_kernel python tools/kernel_bench.py nn 4096000 600 400 dtype=f32
cap gramtc5 gram_tc5_kernel python tools/kernel_bench.py gram 4096000 600 mb=200 dtype=f32
cap resid residual_kernel python tools/kernel_bench.py resid 4096000 300
du -sh gpurun_out
