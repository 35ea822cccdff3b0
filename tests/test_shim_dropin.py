"""Drop-in test: the UNMODIFIED reference library (oracle/_ref_cuda/libxgboost.so) trains through its own
public C API (XGBoosterUpdateOneIter, tree_method=hist, device=cuda) while its "grow_gpu_hist" TreeUpdater
factory is served by our shim (xgboost_b200/lib/libb200hist_xgb_shim.so -> libb200hist.so).  The resulting
model must be identical, tree for tree and bit for bit, to the one the reference's own updater builds.
"""
import ctypes
import json
import os

import numpy as np
import pytest

from tests.conftest import ROOT
from tests.helpers import make_regression

pytestmark = pytest.mark.gpu

SHIM = os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist_xgb_shim.so")


def _train(rc, x, y, params, rounds):
    import torch
    dm = rc.DMatrix.from_cuda(torch.from_numpy(x).cuda(), kind="cuda")
    dm.set_label(y)
    bst = rc.Booster(params, dm)
    for i in range(rounds):
        bst.update(i)
    model = bst.save_json()
    trees = model["learner"]["gradient_booster"]["model"]["trees"]
    keep = ("left_children", "right_children", "split_indices", "split_conditions", "default_left",
            "base_weights", "loss_changes", "sum_hessian")
    return [{k: t[k] for k in keep} for t in trees], bst.predict_margin(dm)


@pytest.mark.parametrize("objective,extra", [("reg:squarederror", {}), ("binary:logistic", {"max_depth": 5}),
                                             ("reg:squarederror", {"grow_policy": "lossguide", "max_leaves": 20, "max_depth": 0})])
def test_reference_booster_trains_identically_through_the_shim(cuda_device, objective, extra):
    from oracle import ref_capi as rc
    if not (rc.available("cuda") and os.path.exists(SHIM)):
        pytest.skip("needs oracle/_ref_cuda/libxgboost.so and the shim (tools/build_shim.sh) -- built where "
                    "/root/reference exists, shipped with the snapshot")
    x, y = make_regression(40000, 24, seed=77, missing=0.05)
    if objective == "binary:logistic":
        y = (y > 0).astype(np.float32)
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": 64, "max_depth": 6, "eta": 0.3,
              "objective": objective, "base_score": 0.5}
    params.update(extra)
    rc.load("cuda")
    shim_loaded = getattr(test_reference_booster_trains_identically_through_the_shim, "_shim", None)
    if shim_loaded is None:
        ref_trees, ref_margin = _train(rc, x, y, params, 4)            # the reference's own updater
        test_reference_booster_trains_identically_through_the_shim._shim = ctypes.CDLL(SHIM, mode=ctypes.RTLD_GLOBAL)
    else:
        ref_trees, ref_margin = _train(rc, x, y, dict(params, updater="grow_gpu_hist_reference"), 4)
    import xgboost_b200 as xb
    before = xb.launch_count()
    our_trees, our_margin = _train(rc, x, y, params, 4)                 # same call, factory now ours
    assert xb.launch_count() > before, "the shim did not route Update() into libb200hist.so"
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))
    # and A/B inside one process through the distinct names
    ab_ref, _ = _train(rc, x, y, dict(params, updater="grow_gpu_hist_reference"), 2)
    ab_our, _ = _train(rc, x, y, dict(params, updater="grow_b200_hist"), 2)
    assert json.dumps(ab_ref) == json.dumps(ab_our)
