#!/bin/bash
# phase-aligned vs sequential walk of the work-list Gram pieces (kernel only), then the gram tests and a short bench
for ph in 0 1; do
  for m in 896 900 600; do
    timeout 120 python tools/kernel_bench.py gram 4096000 $m upper gram_phase=$ph 2>&1 | tail -1 | sed "s/^/phase=$ph /"
  done
  timeout 120 python tools/kernel_bench.py gram 4096000 900 upper distinct gram_phase=$ph 2>&1 | tail -1 | sed "s/^/phase=$ph distinct /"
done
