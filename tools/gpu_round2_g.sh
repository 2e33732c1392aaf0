#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_solver_mid.py tests/test_gpu_full_size.py tests/test_gpu_kernels.py -m gpu -q --maxfail=25 --durations=5 > gpurun_out/pytest_g.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_g.log
tail -14 gpurun_out/pytest_g.log
(
for lp in 40 55 70 85 100 130; do python tools/kernel_bench.py gramcols 4096000 600 300 gram_load_pct=$lp; done
python tools/kernel_bench.py gramcols 4096000 600 300 gram_bk=16
python tools/kernel_bench.py nn 4096000 900 512 nn_warps=16
python tools/kernel_bench.py nn 4096000 900 512
python tools/kernel_bench.py gram 4096000 600 mb=200 dtype=f32
python tools/kernel_bench.py gram 4096000 600 mb=200 dtype=f32 gram_tma=1
python tools/kernel_bench.py gram 4096000 896 upper dtype=f32
python tools/kernel_bench.py gram 4096000 896 upper dtype=f32 gram_tma=1
) > gpurun_out/kb_g.jsonl 2>&1
cat gpurun_out/kb_g.jsonl | cut -c1-400
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_g.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['time_to_solution']['seconds'])
PY
tail -3 gpurun_out/bench_g.err
