#!/bin/bash
# usage: tools/gpu.sh <timeout_s> <script> [gpus]   -- rebuilds every in-tree library first, then ships the tree to the GPU box
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py
T=$1; S=$2; G=${3:-1}
name=$(basename "$S" .sh)
if [ "$G" = "1" ]; then
  /usr/local/graft/bin/gpurun --timeout "$T" -- "bash $S" > "gpurun_out/$name.out" 2>&1 || true
else
  /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "bash $S" > "gpurun_out/$name.out" 2>&1 || true
fi
tail -3 "gpurun_out/$name.out"
