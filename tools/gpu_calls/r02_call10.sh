#!/bin/bash
# Round 2, call 10 (2 GPUs): row split with evaluation by feature slice -- the 2-GPU parity tests (our mirror + the 2-process reference job through the shim), bench N=2 with tree hashes, and the same with the old fused allreduce for A/B
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r02_call10_mgpu.log 2>&1
echo "mgpu rc=$?"; tail -12 gpurun_out/r02_call10_mgpu.log
for mode in sliced full; do
  if [ $mode = full ]; then export B200HIST_ROWSPLIT_FULL=1; else unset B200HIST_ROWSPLIT_FULL; fi
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02_call10_bench_n2_$mode.json 2> gpurun_out/r02_call10_bench_n2_$mode.err
  echo "bench n2 $mode rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_call10_bench_n2_$mode.json').read().strip().splitlines()[-1])
    print("$mode", d["value"], d["ms_per_step"], d["phase_ms"], d["multi_gpu_parity"], (d.get("feature_block_shard") or {}).get("value"))
except Exception as e:
    print("parse failed", e)
PY
  tail -3 gpurun_out/r02_call10_bench_n2_$mode.err
done
