#!/bin/bash
# 1-GPU: full GPU test suite, kernel timings (persistent tall_nn vs one tile per CTA, windowed CSR vs plain), bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 --deselect tests/test_gpu_reftests.py > gpurun_out/pytest_d.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_d.log
tail -25 gpurun_out/pytest_d.log
(
python tools/kernel_bench.py nn 4096000 900 600
python tools/kernel_bench.py nn 4096000 900 600 nn_persist=0
python tools/kernel_bench.py nn 4096000 900 300
python tools/kernel_bench.py nn 4096000 900 300 nn_persist=0
python tools/kernel_bench.py nn 4096000 900 512
python tools/kernel_bench.py nn 4096000 900 512 nn_persist=0
python tools/kernel_bench.py csr 128 128
python tools/kernel_bench.py csr 128 128 csr_window=1
python tools/kernel_bench.py stencil 128 128
) > gpurun_out/kb_d.jsonl 2>&1
cat gpurun_out/kb_d.jsonl
python bench.py --steps 6 --warmup 3 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_d.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['kernels'], d['e2e'], d['time_to_solution']['seconds'], d['gram_cache'], {k:v['ms_per_step'] for k,v in d['windows'].items()})
PY
tail -5 gpurun_out/bench_d.err
