#!/bin/bash
# Round 2, call 5: external-memory tests, agaricus drop-in, quick bench (in-core kernels unchanged?), ncu --set full of the headline hist kernel at 100M
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_extmem.py tests/test_agaricus.py -m gpu -q -p no:cacheprovider > gpurun_out/r02_call5_extmem.log 2>&1
echo "extmem rc=$?"; tail -25 gpurun_out/r02_call5_extmem.log
( time timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e ) > gpurun_out/r02_call5_bench_n1.json 2> gpurun_out/r02_call5_bench_n1.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/r02_call5_bench_n1.json; tail -4 gpurun_out/r02_call5_bench_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:HistKernelPair -s 16 -c 3 -o gpurun_out/r02_hist_pair_100M -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_call5_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/r02_call5_ncu.log; ls -la gpurun_out/*.ncu-rep
