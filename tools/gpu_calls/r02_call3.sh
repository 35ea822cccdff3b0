#!/bin/bash
# Round 2, call 3: ungated drop-in tests with the rebuilt shim; new bench line (N=1); reference arm at full size (how long does it take?)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_next.py tests/test_shim_dropin.py -m gpu -q -p no:cacheprovider > gpurun_out/r02_call3_next.log 2>&1
echo "next+dropin rc=$?"; tail -5 gpurun_out/r02_call3_next.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02_call3_bench_n1.json 2> gpurun_out/r02_call3_bench_n1.err
echo "bench rc=$?"; cat gpurun_out/r02_call3_bench_n1.json; tail -5 gpurun_out/r02_call3_bench_n1.err
( time timeout 1200 python bench.py --impl reference --steps 4 --warmup 1 ) > gpurun_out/r02_call3_ref_full.json 2> gpurun_out/r02_call3_ref_full.err
echo "ref rc=$?"; cat gpurun_out/r02_call3_ref_full.json; tail -5 gpurun_out/r02_call3_ref_full.err
