#!/bin/bash
# second GPU pass (2-GPU box): all GPU tests incl. in-process multi-GPU, then the bench on 2 GPUs through torchrun
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_b.txt
python -m pytest tests -m gpu -q --maxfail=25 --deselect tests/test_gpu_reftests.py --durations=15 > gpurun_out/pytest_b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_b.log
tail -45 gpurun_out/pytest_b.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_b2.json 2> gpurun_out/bench_b2.err
echo "bench2 rc=$?"
tail -c 3000 gpurun_out/bench_b2.json
tail -8 gpurun_out/bench_b2.err
