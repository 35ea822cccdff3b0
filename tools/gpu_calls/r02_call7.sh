#!/bin/bash
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/p tools/hist_pipe_probe.cu -lcuda > gpurun_out/r02_pipe_probe_build.log 2>&1 && timeout 300 /tmp/p > gpurun_out/r02_pipe_probe.log 2>&1
echo "probe rc=$?"; cat gpurun_out/r02_pipe_probe.log
