"""xgboost_b200 -- Blackwell-native (sm_100a) `hist` tree updater behind the reference's
TreeUpdater / EllpackPage interface.  Only the hot path lives here:

  csrc/          hand-written CUDA kernels + the C ABI (include/b200hist.h) -> lib/libb200hist.so
  csrc/shim/     the reference-side binding (xgboost::TreeUpdater) -> lib/libb200hist_xgb_shim.so
  lib.py         ctypes binding of the C ABI
  updater.py     host-side mirror of the reference interface (HistogramCuts, QuantileSketch, EllpackPage,
                 GPUHistMaker incl. sampling / constraints / peer exchange, RegTree, Forest = prediction on pages)
  components.py  per-stage entry points used by the parity tests, measured shared-atomic ceiling

There is no CPU fallback: without the built library or without a CUDA device every compute call
raises B200HistError.
"""
from .lib import B200HistError, TrainParam, launch_count  # noqa: F401
from .updater import (Communicator, EllpackPage, Forest, GPUHistMaker, HistogramCuts, QuantileSketch, RegTree,  # noqa: F401
                      shard_groups)

__all__ = ["B200HistError", "TrainParam", "EllpackPage", "GPUHistMaker", "HistogramCuts", "QuantileSketch", "RegTree",
           "Forest", "Communicator", "shard_groups", "launch_count"]
