#!/bin/bash
# Round 2, first GPU call: gather4 throughput probe, reference fixtures that wait for a GPU, the 84 gated tests (no -x).
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_call1_env.log; nproc >> gpurun_out/r02_call1_env.log; free -g >> gpurun_out/r02_call1_env.log
timeout 200 nvcc -O3 -lineinfo -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/gather4_throughput \
  tools/gather4_throughput.cu -lcuda > gpurun_out/gather4_build.log 2>&1 \
  && timeout 120 /tmp/gather4_throughput > gpurun_out/gather4_throughput.log 2>&1
echo "gather4 rc=$?"
timeout 300 python tools/make_golden_gpu_next.py --out gpurun_out/ref_gpu_next_trees.json > gpurun_out/next_fixture.log 2>&1
echo "fixture rc=$?"
B200HIST_RUN_UNVALIDATED=1 timeout 1200 python -m pytest tests/test_gpu_next.py -m gpu -q -p no:cacheprovider \
  > gpurun_out/next_tests.log 2>&1
echo "gated tests rc=$?" | tee -a gpurun_out/next_tests.log
tail -40 gpurun_out/gather4_throughput.log 2>/dev/null
tail -60 gpurun_out/next_tests.log
