#!/bin/bash
# Round 2, call 16: what the driver runs at round end -- pytest -m gpu, smoke, bench (our arm and the reference arm, driver's flags) -- plus the ncu launch list of the tree kernels
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r02_call16_suite.log 2>&1
echo "suite rc=$?"; tail -5 gpurun_out/r02_call16_suite.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02_call16_ref.json 2> gpurun_out/r02_call16_ref.err
echo "ref rc=$?"; cut -c1-900 gpurun_out/r02_call16_ref.json; tail -4 gpurun_out/r02_call16_ref.err
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02_call16_bench.json 2> gpurun_out/r02_call16_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_call16_bench.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","reference_gpu","other_configs","gpu_launches","clocks"):
    print(k, json.dumps(d.get(k))[:1200])
PY
tail -4 gpurun_out/r02_call16_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_tree_launches.csv \
  -k 'regex:^(HistKernel|PartitionCount|PartitionScatter|ScanKernel|ApplyKernel|Finalise|FinalList|Quant|Subtract|FeatureScan|NodeReduce|PredCache|InitRoot)' \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_call16_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r02_tree_launches.csv
