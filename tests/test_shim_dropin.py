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


def _ensure_shim(rc):
    rc.load("cuda")
    f = test_reference_booster_trains_identically_through_the_shim
    if getattr(f, "_shim", None) is None:
        f._shim = ctypes.CDLL(SHIM, mode=ctypes.RTLD_GLOBAL)


@pytest.mark.parametrize("extra", [{"subsample": 0.7}, {"colsample_bytree": 0.5}, {"subsample": 0.5, "colsample_bytree": 0.6},
                                   {"colsample_bylevel": 0.5},
                                   {"colsample_bylevel": 0.6, "colsample_bytree": 0.8, "subsample": 0.7, "max_depth": 8},
                                   {"colsample_bylevel": 0.5, "grow_policy": "lossguide", "max_leaves": 20, "max_depth": 0},
                                   {"monotone_constraints": "(1,-1,0,1,0,0,-1)"},
                                   {"monotone_constraints": "(1,1,1,1)", "subsample": 0.8, "max_delta_step": 0.7},
                                   {"subsample": 0.8, "seed": 123, "grow_policy": "lossguide", "max_leaves": 24, "max_depth": 0}],
                         ids=lambda e: "_".join("%s%s" % kv for kv in e.items()))
def test_sampling_through_the_shim_matches_reference(cuda_device, extra):
    """Row sampling (uniform) and colsample_bytree: the shim draws from the Booster's RNG exactly where the reference
    does (sampler seed, then the column sampler's shuffle seed, once per tree); the row trial itself
    (minstd_rand + discard + uniform float) runs in our quantise kernel.  Models must match bit for bit."""
    from oracle import ref_capi as rc
    if not (rc.available("cuda") and os.path.exists(SHIM)):
        pytest.skip("needs oracle/_ref_cuda/libxgboost.so and the shim")
    _ensure_shim(rc)
    x, y = make_regression(30000, 40, seed=78)
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": 64, "max_depth": 5, "eta": 0.3,
              "objective": "reg:squarederror", "base_score": 0.5}
    params.update(extra)
    ref_trees, ref_margin = _train(rc, x, y, dict(params, updater="grow_gpu_hist_reference"), 4)
    our_trees, our_margin = _train(rc, x, y, dict(params, updater="grow_b200_hist"), 4)
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))
    # sampling really happened: the unsampled model differs
    plain, _ = _train(rc, x, y, dict({k: v for k, v in params.items()
                                      if k not in ("subsample", "colsample_bytree", "colsample_bylevel",
                                                   "monotone_constraints")},
                                     updater="grow_b200_hist"), 4)
    assert json.dumps(plain) != json.dumps(our_trees)


def test_sparse_dmatrix_through_the_shim_matches_reference(cuda_device):
    """A CSR DMatrix becomes a sparse ELLPACK page in the reference (row_stride < n_features); the shim imports it
    into our sparse page (uint32 global-bin symbols) and must reproduce the reference's trees."""
    from oracle import ref_capi as rc
    from tests.helpers import make_sparse_regression
    if not (rc.available("cuda") and os.path.exists(SHIM)):
        pytest.skip("needs oracle/_ref_cuda/libxgboost.so and the shim")
    _ensure_shim(rc)
    _, y, indptr, indices, values = make_sparse_regression(20000, 200, 0.04, 79)

    def train(updater):
        dm = rc.DMatrix.from_csr(indptr, indices, values, 200, kind="cuda")
        dm.set_label(y)
        bst = rc.Booster({"tree_method": "hist", "device": "cuda:0", "max_bin": 64, "max_depth": 6, "eta": 0.3,
                          "objective": "reg:squarederror", "base_score": 0.5, "updater": updater}, dm)
        for i in range(3):
            bst.update(i)
        trees = bst.save_json()["learner"]["gradient_booster"]["model"]["trees"]
        keep = ("left_children", "right_children", "split_indices", "split_conditions", "default_left",
                "base_weights", "loss_changes", "sum_hessian")
        return [{k: t[k] for k in keep} for t in trees], bst.predict_margin(dm)

    ref_trees, ref_margin = train("grow_gpu_hist_reference")
    our_trees, our_margin = train("grow_b200_hist")
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))


def test_external_memory_dmatrix_through_the_shim_matches_reference(cuda_device):
    """ExtMemQuantileDMatrix (four batches -> four Ellpack pages in a host cache): the reference streams the pages
    through its updater; the shim decodes each page into its row range of one resident page.  Same trees."""
    import torch
    from oracle import ref_capi as rc
    if not (rc.available("cuda") and os.path.exists(SHIM)):
        pytest.skip("needs oracle/_ref_cuda/libxgboost.so and the shim")
    _ensure_shim(rc)
    x, y = make_regression(48000, 20, seed=80, missing=0.03)
    parts = [x[:10000], x[10000:31000], x[31000:31007], x[31007:]]

    def train(updater):
        dm = rc.DMatrix.extmem_quantile_from_batches([torch.from_numpy(p).cuda() for p in parts], max_bin=64, kind="cuda")
        dm.set_label(y)
        bst = rc.Booster({"tree_method": "hist", "device": "cuda:0", "max_bin": 64, "max_depth": 6, "eta": 0.3,
                          "objective": "reg:squarederror", "base_score": 0.5, "updater": updater}, dm)
        for i in range(3):
            bst.update(i)
        trees = bst.save_json()["learner"]["gradient_booster"]["model"]["trees"]
        keep = ("left_children", "right_children", "split_indices", "split_conditions", "default_left",
                "base_weights", "loss_changes", "sum_hessian")
        return [{k: t[k] for k in keep} for t in trees], bst.predict_margin(dm)

    ref_trees, ref_margin = train("grow_gpu_hist_reference")
    import xgboost_b200 as xb
    before = xb.launch_count()
    our_trees, our_margin = train("grow_b200_hist")
    assert xb.launch_count() > before
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))


@pytest.mark.parametrize("extra", [{"objective": "reg:absoluteerror"},
                                   {"objective": "reg:quantileerror", "quantile_alpha": 0.3},
                                   {"objective": "reg:absoluteerror", "subsample": 0.7, "max_depth": 4}],
                         ids=["absoluteerror", "quantileerror", "absoluteerror_subsample"])
def test_adaptive_objectives_through_the_shim_match_reference(cuda_device, extra):
    """reg:absoluteerror / reg:quantileerror rewrite the tree's leaves AFTER Update (ObjFunction::UpdateTreeLeaf,
    gbtree.cc:205-222) from the positions the updater returned, and UpdatePredictionCache must then add the REWRITTEN
    leaves (the reference reads p_tree->GetNodes() there, updater_gpu_hist.cu:594-598).  From the second iteration on
    the gradients depend on that cache: three rounds, identical models and margins."""
    from oracle import ref_capi as rc
    if not (rc.available("cuda") and os.path.exists(SHIM)):
        pytest.skip("needs oracle/_ref_cuda/libxgboost.so and the shim")
    _ensure_shim(rc)
    x, y = make_regression(30000, 16, seed=83, missing=0.02)
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": 64, "max_depth": 5, "eta": 0.3, "base_score": 0.5}
    params.update(extra)
    ref_trees, ref_margin = _train(rc, x, y, dict(params, updater="grow_gpu_hist_reference"), 3)
    our_trees, our_margin = _train(rc, x, y, dict(params, updater="grow_b200_hist"), 3)
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))


def test_small_histogram_cache_through_the_shim_matches_reference(cuda_device):
    """max_cached_hist_node (HistMakerTrainParam) below the tree's size: the reference overflows its cache and rebuilds
    what it cannot subtract (updater_gpu_hist.cu:309-325); ours wraps its ring and builds both children.  Same model."""
    from oracle import ref_capi as rc
    if not (rc.available("cuda") and os.path.exists(SHIM)):
        pytest.skip("needs oracle/_ref_cuda/libxgboost.so and the shim")
    _ensure_shim(rc)
    x, y = make_regression(40000, 12, seed=229)
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": 32, "max_depth": 8, "eta": 0.3,
              "objective": "reg:squarederror", "base_score": 0.5, "max_cached_hist_node": 24}
    ref_trees, ref_margin = _train(rc, x, y, dict(params, updater="grow_gpu_hist_reference"), 3)
    our_trees, our_margin = _train(rc, x, y, dict(params, updater="grow_b200_hist"), 3)
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))
