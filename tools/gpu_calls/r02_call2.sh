#!/bin/bash
# Round 2, call 2: the ungated suite + the new drop-in tests (multi-class stride, adaptive objectives) + parity at BASELINE sizes.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_next.py tests/test_shim_dropin.py -m gpu -q -p no:cacheprovider > gpurun_out/r02_call2_next.log 2>&1
echo "next+dropin rc=$?"; tail -15 gpurun_out/r02_call2_next.log
timeout 1200 python -m pytest tests/test_shim_scale.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02_call2_scale.log 2>&1
echo "scale rc=$?"; tail -30 gpurun_out/r02_call2_scale.log
