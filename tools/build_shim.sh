#!/bin/bash
# Builds the reference-side bindings (xgboost::TreeUpdater shim + the gated xgboost::Predictor shim) against the reference's own headers.
# Needs /root/reference (this container) + oracle/_ref_cuda/libxgboost.so + xgboost_b200/lib/libb200hist.so.
# Output: xgboost_b200/lib/libb200hist_xgb_shim.so (git-ignored; travels to the GPU box with the snapshot).
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${REF:-/root/reference}
OUT="$ROOT/xgboost_b200/lib/libb200hist_xgb_shim.so"
/usr/local/cuda/bin/nvcc -std=c++17 -O2 -shared -ccbin /usr/bin/g++ -Xcompiler=-fPIC -Xcompiler=-fopenmp -w \
  --expt-extended-lambda --expt-relaxed-constexpr --default-stream per-thread \
  -gencode arch=compute_100,code=sm_100 \
  -DNDEBUG -DDMLC_LOG_CUSTOMIZE=1 -D_MWAITXINTRIN_H_INCLUDED -DDMLC_USE_CXX11=1 -DDMLC_USE_CXX14=1 -DXGBOOST_USE_CUDA=1 \
  -I"$REF/include" -I"$REF/dmlc-core/include" -I"$REF/gputreeshap" -I"$REF" \
  -o "$OUT" "$ROOT/xgboost_b200/csrc/shim/b200_tree_updater.cu" "$ROOT/xgboost_b200/csrc/shim/b200_predictor.cu" \
  -L"$ROOT/oracle/_ref_cuda" -lxgboost -L"$ROOT/xgboost_b200/lib" -lb200hist \
  -Xlinker -rpath -Xlinker '$ORIGIN' -Xlinker -rpath -Xlinker '$ORIGIN/../../oracle/_ref_cuda'
echo "built $OUT"
