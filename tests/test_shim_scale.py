"""Drop-in parity at BASELINE.json's sizes: ONE QuantileDMatrix of the unmodified reference (oracle/_ref_cuda), two
Boosters on it -- the reference's own updater ("grow_gpu_hist_reference") and ours through the shim
("grow_b200_hist") -- and the models must be the same JSON, the margins the same bits.

  C2        10M x 64 dense, 256 bins, depth 6           (2 trees: the second one sees non-trivial gradients)
  C3        100M x 256 dense, 256 bins, depth 8         (the headline matrix, 1 tree)
  C5-shape  5M x 1024, 1 % present, 256 bins, depth 10  (sparse ELLPACK page, 1 tree)

This is where 32-bit row x stride products, > 2^24-row windows and 195k-entry tile tables would break; the small
twins in tests/test_shim_dropin.py cannot see them.  Batches of 1M rows are generated on the device (Philox, seed
2024 + batch -- SURVEY.md 8(d)) and never exist at once.  Reference test this mirrors:
tests/cpp/tree/test_gpu_hist.cu:24-48, tests/python-gpu/test_gpu_updaters.py:37-50.
"""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu

SHIM = os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist_xgb_shim.so")
KEEP = ("left_children", "right_children", "parents", "split_indices", "split_conditions", "default_left",
        "base_weights", "loss_changes", "sum_hessian")


def _rc_or_skip():
    from oracle import ref_capi as rc
    from tests import test_shim_dropin as sd
    if not (rc.available("cuda") and os.path.exists(SHIM)):
        pytest.skip("needs oracle/_ref_cuda/libxgboost.so and the shim (built where /root/reference exists)")
    rc.load("cuda")
    f = sd.test_reference_booster_trains_identically_through_the_shim
    if getattr(f, "_shim", None) is None:
        f._shim = ctypes.CDLL(SHIM, mode=ctypes.RTLD_GLOBAL)
    return rc


def _batch(torch, dev, f, rows, b, density):
    m = min(1_000_000, rows - b * 1_000_000)
    g = torch.Generator(device=dev).manual_seed(2024 + b)
    x = torch.randn((m, f), generator=g, device=dev)
    if density < 1.0:
        x = torch.where(torch.rand((m, f), generator=g, device=dev) < density, x, torch.full_like(x, float("nan")))
    return x, g


def _labels(torch, dev, f, rows, density):
    w = torch.randn(f, generator=torch.Generator(device=dev).manual_seed(7), device=dev)
    y = torch.empty(rows, dtype=torch.float32, device=dev)
    nb = (rows + 999_999) // 1_000_000
    for b in range(nb):
        x, g = _batch(torch, dev, f, rows, b, density)
        y[b * 1_000_000:b * 1_000_000 + x.shape[0]] = torch.nan_to_num(x) @ w + \
            0.1 * torch.randn(x.shape[0], generator=g, device=dev)
        del x
    return y.cpu().numpy()


def _train(rc, dm, params, updater, rounds):
    bst = rc.Booster(dict(params, updater=updater), dm)
    for i in range(rounds):
        bst.update(i)
    trees = bst.save_json()["learner"]["gradient_booster"]["model"]["trees"]
    margin = bst.predict_margin(dm)
    del bst
    return [{k: t[k] for k in KEEP} for t in trees], margin


@pytest.mark.parametrize("rows,f,depth,density,rounds", [
    (10_000_000, 64, 6, 1.0, 2),
    (100_000_000, 256, 8, 1.0, 1),
    (5_000_000, 1024, 10, 0.01, 1),
], ids=["C2_10Mx64_d6", "C3_100Mx256_d8", "C5shape_5Mx1024_1pct_d10"])
def test_one_tree_at_baseline_size_is_bit_identical_to_the_reference(cuda_device, rows, f, depth, density, rounds):
    import torch
    import xgboost_b200 as xb
    rc = _rc_or_skip()
    dev = cuda_device
    free, _ = torch.cuda.mem_get_info(dev)
    need = rows * f * (3.2 if density == 1.0 else 0.5) + 8e9   # reference page + ours (+ column-major copy) + state
    if free < need:
        pytest.skip("needs %.0f GB of free HBM" % (need / 1e9))
    nb = (rows + 999_999) // 1_000_000
    batches = rc.LazyBatches(nb, lambda b: _batch(torch, dev, f, rows, b, density)[0])
    dm = rc.DMatrix.quantile_from_batches(batches, max_bin=256, kind="cuda")
    dm.set_label(_labels(torch, dev, f, rows, density))
    torch.cuda.empty_cache()
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": 256, "max_depth": depth, "eta": 0.3,
              "objective": "reg:squarederror", "base_score": 0.5, "lambda": 1, "min_child_weight": 1}
    ref_trees, ref_margin = _train(rc, dm, params, "grow_gpu_hist_reference", rounds)
    before = xb.launch_count()
    our_trees, our_margin = _train(rc, dm, params, "grow_b200_hist", rounds)
    assert xb.launch_count() > before, "the shim did not route Update() into libb200hist.so"
    assert len(ref_trees) == rounds and len(ref_trees[-1]["left_children"]) > 2 ** (depth - 2), "degenerate tree"
    for t, (a, b) in enumerate(zip(our_trees, ref_trees)):
        for k in KEEP:
            assert a[k] == b[k], "tree %d: %s differs" % (t, k)
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))
    print("parity %dx%d depth %d: %d nodes, model sha %s" %
          (rows, f, depth, len(ref_trees[-1]["left_children"]),
           hashlib.sha256(json.dumps(our_trees).encode()).hexdigest()[:16]))
    del dm
    torch.cuda.empty_cache()
