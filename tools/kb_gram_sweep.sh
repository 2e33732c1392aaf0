python -m pytest tests/test_gpu_kernels.py -x -q -k gram 2>&1 | tail -3
for m in 128 256 300 384 450 512 600 750 896 900 928 960 1024 1152; do
  python tools/kernel_bench.py gram 4096000 $m upper
  python tools/kernel_bench.py gram 4096000 $m upper gram_wl=0
done 2>&1 | tee gpurun_out/kb_gram7.log
for a in "4096000 900 upper gram_strip_max=-1 gram_load_pct=100" "4096000 928 upper gram_strip_max=-1" "4096000 960 upper gram_strip_max=100" "512000 900 upper" "512000 900 upper gram_wl=0" "2097152 384 upper" "2097152 384 upper gram_wl=0"; do python tools/kernel_bench.py gram $a; done 2>&1 | tee -a gpurun_out/kb_gram7.log
