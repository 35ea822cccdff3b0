#!/bin/bash
# Round 2, call 12 (2 GPUs): deferred read-back test, bench N=1 and N=2 with it
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "deferred" > gpurun_out/r02_call12_async.log 2>&1
echo "async rc=$?"; tail -6 gpurun_out/r02_call12_async.log
( time timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline ) > gpurun_out/r02_call12_bench_n1.json 2> gpurun_out/r02_call12_bench_n1.err
echo "bench n1 rc=$?"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02_call12_bench_n2.json 2> gpurun_out/r02_call12_bench_n2.err
echo "bench n2 rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r02_call12_bench_n1.json","gpurun_out/r02_call12_bench_n2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["step_parts_ms"], d["multi_gpu_parity"], d["roofline"]["frac"])
    except Exception as e:
        print("parse failed", f, e)
PY
tail -3 gpurun_out/r02_call12_bench_n1.err gpurun_out/r02_call12_bench_n2.err
