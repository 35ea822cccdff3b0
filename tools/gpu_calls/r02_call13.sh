#!/bin/bash
# Round 2, call 13: the whole GPU suite, the full bench line (e2e, reference_gpu, C2, C5, cpu_baseline), the ncu launch list of the bench command
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_call13_suite.log 2>&1
echo "suite rc=$?"; tail -6 gpurun_out/r02_call13_suite.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02_call13_bench_n1.json 2> gpurun_out/r02_call13_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_call13_bench_n1.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","reference_gpu","cpu_baseline","other_configs","roofline"):
    print(k, json.dumps(d.get(k))[:900])
PY
tail -4 gpurun_out/r02_call13_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_call13_ncu_launches.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/r02_launches.csv
