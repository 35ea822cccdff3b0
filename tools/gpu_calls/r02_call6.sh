#!/bin/bash
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/probe tools/smem_layout_probe.cu > gpurun_out/r02_layout_probe_build.log 2>&1 && timeout 120 /tmp/probe > gpurun_out/r02_layout_probe.log 2>&1
echo "probe rc=$?"; cat gpurun_out/r02_layout_probe.log
