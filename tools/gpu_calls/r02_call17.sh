#!/bin/bash
# Round 2, call 17 (2 GPUs): the 2-GPU tests, then the driver's N=2 commands (reference arm under torchrun, our arm with the distributed e2e)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r02_call17_mgpu.log 2>&1
echo "mgpu rc=$?"; tail -4 gpurun_out/r02_call17_mgpu.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02_call17_bench_n2.json 2> gpurun_out/r02_call17_bench_n2.err
echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_call17_bench_n2.json').read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","e2e","reference_gpu","multi_gpu_parity","feature_block_shard","cpu_baseline","step_parts_ms"):
        print(k, json.dumps(d.get(k))[:700])
except Exception as e:
    print("parse failed", e)
PY
grep -v "WARNING\|^$" gpurun_out/r02_call17_bench_n2.err | tail -12
