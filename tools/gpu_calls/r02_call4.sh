#!/bin/bash
# Round 2, call 4 (2 GPUs): the whole GPU suite incl. the 2-GPU parity tests and the 2-process reference job through the shim;
# bench at N=2 (tree hashes); reference arm under torchrun (thread clamp fixed?)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi -L > gpurun_out/r02_call4_env.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_call4_suite.log 2>&1
echo "suite rc=$?"; tail -15 gpurun_out/r02_call4_suite.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r02_call4_bench_n2.json 2> gpurun_out/r02_call4_bench_n2.err
echo "bench n2 rc=$?"; cat gpurun_out/r02_call4_bench_n2.json | cut -c1-3000; tail -5 gpurun_out/r02_call4_bench_n2.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) > gpurun_out/r02_call4_ref_n2.json 2> gpurun_out/r02_call4_ref_n2.err
echo "ref n2 rc=$?"; cat gpurun_out/r02_call4_ref_n2.json | cut -c1-2000; tail -5 gpurun_out/r02_call4_ref_n2.err
