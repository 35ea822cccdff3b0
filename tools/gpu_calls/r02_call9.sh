#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "node_batches or whole_tree or tree_matches" > gpurun_out/r02_call9_batch.log 2>&1
echo "batch rc=$?"; tail -12 gpurun_out/r02_call9_batch.log
timeout 900 python tools/bench_predictor.py --rows 10000000 --features 64 --trees 50 --depth 6 > gpurun_out/r02_predictor_bench.json 2> gpurun_out/r02_predictor_bench.err
echo "predictor rc=$?"; cat gpurun_out/r02_predictor_bench.json; tail -3 gpurun_out/r02_predictor_bench.err
