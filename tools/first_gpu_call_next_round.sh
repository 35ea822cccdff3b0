#!/bin/bash
# The first GPU call of the next round, in one piece (gpurun --timeout 1500 -- 'bash tools/first_gpu_call_next_round.sh'):
#   1. the gated tests of the paths written late in round 1 (no -x: the log must show everything that fails),
#   2. the default GPU suite (what the driver runs),
#   3. a short bench line.
# Logs land in gpurun_out/ and come back with the call.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
# 0. fixtures of the unmodified reference for the oracle restatements that still wait for a GPU pin (seconds);
#    copy gpurun_out/ref_gpu_next_trees.json to tests/golden/ afterwards: two waiting CPU tests wake up
timeout 300 python tools/make_golden_gpu_next.py --out gpurun_out/ref_gpu_next_trees.json > gpurun_out/next_fixture.log 2>&1
echo "fixture rc=$?"
B200HIST_RUN_UNVALIDATED=1 timeout 1200 python -m pytest tests/test_gpu_next.py -m gpu -q -p no:cacheprovider \
  > gpurun_out/next_tests.log 2>&1
echo "gated tests rc=$?" | tee -a gpurun_out/next_tests.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_suite.log 2>&1
echo "default suite rc=$?" | tee -a gpurun_out/gpu_suite.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_line.json 2> gpurun_out/bench.err
echo "bench rc=$?"
tail -5 gpurun_out/next_tests.log gpurun_out/gpu_suite.log
cat gpurun_out/bench_line.json
# 4. DESIGN (f) rank 1: can TMA gather4 carry the histogram kernel's row gathers?  (compiles here: the box has nvcc;
#    cross-compiled clean in round 1, never run -- its own timeout keeps a wedged copy from costing the call)
timeout 200 nvcc -O3 -lineinfo -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/gather4_throughput \
  tools/gather4_throughput.cu -lcuda > gpurun_out/gather4_build.log 2>&1 \
  && timeout 120 /tmp/gather4_throughput > gpurun_out/gather4_throughput.log 2>&1
echo "gather4 rc=$?"
tail -30 gpurun_out/gather4_throughput.log 2>/dev/null
