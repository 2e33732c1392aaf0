#!/bin/bash
# unpreconditioned C5 solve to convergence with the default path (int8 tensor kernels), and the C2-C4 configurations
mkdir -p gpurun_out
timeout 900 python tools/full_solve.py 160 150 1e-8 2000 > gpurun_out/full_solve_c5_r02.log 2>&1
tail -2 gpurun_out/full_solve_c5_r02.log | cut -c1-900
