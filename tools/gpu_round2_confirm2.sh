#!/bin/bash
# 2-GPU confirmation of the round's last build: in-process multi-GPU tests + the torchrun bench at N = 2
mkdir -p gpurun_out
SECONDS=0
timeout 300 python -m pytest tests/test_gpu_multi_inprocess.py -m gpu -q > gpurun_out/pytest_confirm_2gpu.log 2>&1
echo "pytest rc=$? wall ${SECONDS}s" | tee -a gpurun_out/pytest_confirm_2gpu.log
tail -5 gpurun_out/pytest_confirm_2gpu.log
T0=$SECONDS
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_confirm_2gpu.json 2> gpurun_out/bench_confirm_2gpu.err
echo "bench rc=$? wall $((SECONDS-T0))s"; tail -3 gpurun_out/bench_confirm_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_confirm_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches','n_gpus')}, d['e2e'], d['roofline']['frac'], d['clocks'])
print({k:v.get('ms_per_step') for k,v in d['kernels'].items()}, {k:v['ms_per_step'] for k,v in d['windows'].items()}, d['time_to_solution'])
PY
