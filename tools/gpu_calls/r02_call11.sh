#!/bin/bash
# Round 2, call 11 (8 GPUs): bench at N=8 (tree hashes vs 1 GPU, sliced evaluation), then the fused allreduce for A/B if time allows
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi -L | wc -l
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02_call11_bench_n8.json 2> gpurun_out/r02_call11_bench_n8.err
echo "bench n8 rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_call11_bench_n8.json').read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["phase_ms"], d["step_parts_ms"], d["multi_gpu_parity"], d.get("feature_block_shard"))
except Exception as e:
    print("parse failed", e)
PY
tail -5 gpurun_out/r02_call11_bench_n8.err
export B200HIST_ROWSPLIT_FULL=1 B200_BENCH_ALSO_FEATURES=0
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02_call11_bench_n8_full.json 2> gpurun_out/r02_call11_bench_n8_full.err
echo "bench n8 full rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_call11_bench_n8_full.json').read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["phase_ms"])
except Exception as e:
    print("parse failed", e)
PY
