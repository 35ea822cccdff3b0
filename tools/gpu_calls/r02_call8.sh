#!/bin/bash
# Round 2, call 8: whole GPU suite with the predictor binding on by default; predictor timing; bench with the C2 side measurement
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_call8_suite.log 2>&1
echo "suite rc=$?"; tail -8 gpurun_out/r02_call8_suite.log
timeout 900 python tools/bench_predictor.py --rows 10000000 --features 64 --trees 50 --depth 6 > gpurun_out/r02_predictor_bench.json 2> gpurun_out/r02_predictor_bench.err
echo "predictor rc=$?"; cat gpurun_out/r02_predictor_bench.json; tail -3 gpurun_out/r02_predictor_bench.err
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02_call8_bench_n1.json 2> gpurun_out/r02_call8_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_call8_bench_n1.json').read().strip().splitlines()[0])
for k in ("value","ms_per_step","e2e","reference_gpu","cpu_baseline","other_configs"):
    print(k, json.dumps(d.get(k))[:700])
PY
tail -4 gpurun_out/r02_call8_bench_n1.err
