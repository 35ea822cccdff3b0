"""Run in its own process by tests/test_gpu_next.py (the shim reads B200HIST_SHIM_PREDICTOR when it is loaded).

Trains the UNMODIFIED reference (oracle/_ref_cuda/libxgboost.so) with its own updater on a QuantileDMatrix, predicts a
second QuantileDMatrix that has its OWN cuts (so split conditions are not cuts of the predicted page), and writes the
margins plus the number of libb200hist kernel launches the prediction caused.  With B200HIST_SHIM_PREDICTOR=1 the
reference's "gpu_predictor" factory is served by xgboost_b200/csrc/shim/b200_predictor.cu; without it the run is the
reference's own.  usage: python -m tests.shim_predictor_check <out.npz> <case>"""
import ctypes
import os
import sys

import numpy as np


def main(out_path: str, case: str) -> None:
    import torch
    from oracle import ref_capi as rc
    from tests.conftest import ROOT
    from tests.helpers import make_regression
    rc.load("cuda")
    core = ctypes.CDLL(os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist.so"), mode=ctypes.RTLD_GLOBAL)
    shim = ctypes.CDLL(os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist_xgb_shim.so"), mode=ctypes.RTLD_GLOBAL)
    core.b200hist_launch_count.restype = ctypes.c_uint64
    missing = {"dense": 0.0, "missing": 0.08, "wide_bins": 0.03, "multiclass": 0.04}[case]
    max_bin = {"dense": 64, "missing": 256, "wide_bins": 600, "multiclass": 64}[case]
    x, y = make_regression(40000, 14, seed=71, missing=missing)
    xv, _ = make_regression(15000, 14, seed=72, missing=missing)
    dtrain = rc.DMatrix.quantile_from_batches([torch.from_numpy(x).cuda()], max_bin=max_bin, kind="cuda")
    dtrain.set_label(y)
    dvalid = rc.DMatrix.quantile_from_batches([torch.from_numpy(xv).cuda()], max_bin=max(16, max_bin // 3), kind="cuda")
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": max_bin, "max_depth": 6, "eta": 0.3,
              "objective": "reg:squarederror", "base_score": 0.5, "updater": "grow_gpu_hist_reference"}  # the reference's own trees
    if case == "multiclass":   # tree_info names each tree's output; leaves go straight onto the output
        y = np.digitize(y, np.quantile(y, [0.25, 0.5, 0.75])).astype(np.float32)
        dtrain.set_label(y)
        params.update(objective="multi:softprob", num_class=4)
    bst = rc.Booster(params, dtrain)
    for i in range(6):
        bst.update(i)
    before = int(core.b200hist_launch_count())
    valid = bst.predict_margin(dvalid)
    train = bst.predict_margin(dtrain)
    leaf = bst.predict_margin(dvalid, pred_type=6)
    launches = int(core.b200hist_launch_count()) - before
    np.savez(out_path, valid=valid, train=train, leaf=leaf, launches=np.int64(launches))
    del shim


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
