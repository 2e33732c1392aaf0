#!/bin/bash
# 2 GPUs: in-process multi-GPU tests of the reference entry points, 2-rank bench (device and e2e legs), 2-rank distributed parity check
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_inprocess.py -m gpu -q --durations=5 > gpurun_out/pytest_q.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_q.log
tail -8 gpurun_out/pytest_q.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 12 --warmup 4 > gpurun_out/bench_q2.json 2> gpurun_out/bench_q2.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_q2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['kernels'], d['e2e'], d['time_to_solution'].get('seconds'), {k:v['ms_per_step'] for k,v in d['windows'].items()})
PY
