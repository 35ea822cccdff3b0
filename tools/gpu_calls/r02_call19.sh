#!/bin/bash
# Round 2, call 19 (8 GPUs): the driver's N=8 command (our arm incl. the distributed e2e through the reference job)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r02_call19_bench_n8.json 2> gpurun_out/r02_call19_bench_n8.err
echo "bench n8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_call19_bench_n8.json').read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","e2e","reference_gpu","multi_gpu_parity","feature_block_shard","phase_ms","step_parts_ms"):
        print(k, json.dumps(d.get(k))[:600])
except Exception as e:
    print("parse failed", e)
PY
grep -v "WARNING\|^$\|\*\*\*\*\|OMP_NUM" gpurun_out/r02_call19_bench_n8.err | tail -8
