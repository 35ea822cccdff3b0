#!/usr/bin/env python
"""One short GPU call (no torch import): trees of the UNMODIFIED reference gpu_hist (oracle/_ref_cuda) for the paths
whose oracle restatements are not yet pinned bit for bit on a GPU reference:

  * categorical features (so far pinned structurally on the reference's CPU hist trees),
  * gradient-based row sampling (so far property-tested only),
  * colsample_bynode (the oracle regenerates the reference's per-node draws from the Booster's mt19937 through its
    restatement of thrust::shuffle, which is pinned on the toolkit's own vectors; the tree-level order of the draws
    is what this fixture pins).

    python tools/make_golden_gpu_next.py --out gpurun_out/ref_gpu_next_trees.json
    cp gpurun_out/ref_gpu_next_trees.json tests/golden/        # activates the tests that wait for it

tests/test_reference_golden.py::test_oracle_reproduces_reference_gpu_{categorical_trees,gradient_based_sampling} and ..._colsample_bynode skip
while the fixture is absent.  Inputs regenerate from the seeds through tests/helpers.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_capi as rc  # noqa: E402
from tests.helpers import make_categorical_data, make_regression  # noqa: E402

CASES = [
    dict(name="cat_onehot4", kind="categorical", seed=311, n=8000, max_bin=64, rounds=2,
         params=dict(max_depth=5, eta=0.5, max_cat_to_onehot=4, max_cat_threshold=64)),
    dict(name="cat_all_partition", kind="categorical", seed=312, n=8000, max_bin=64, rounds=2,
         params=dict(max_depth=5, eta=0.5, max_cat_to_onehot=1, max_cat_threshold=64)),
    dict(name="cat_all_onehot", kind="categorical", seed=313, n=8000, max_bin=64, rounds=2,
         params=dict(max_depth=5, eta=0.5, max_cat_to_onehot=100, max_cat_threshold=64)),
    dict(name="cat_threshold8_lossguide", kind="categorical", seed=314, n=8000, max_bin=64, rounds=2,
         params=dict(max_depth=0, max_leaves=20, grow_policy="lossguide", eta=0.5, max_cat_to_onehot=4,
                     max_cat_threshold=8)),
    dict(name="bynode_05", kind="bynode", seed=331, n=20000, f=24, max_bin=64, rounds=3,
         params=dict(max_depth=5, eta=0.3, colsample_bynode=0.5)),
    dict(name="bynode_06_bytree_07_lossguide", kind="bynode", seed=332, n=20000, f=24, max_bin=64, rounds=3,
         params=dict(max_depth=0, max_leaves=20, grow_policy="lossguide", eta=0.3, colsample_bynode=0.6,
                     colsample_bytree=0.7, seed=3)),
    dict(name="bytree_05", kind="colsample", seed=341, n=20000, f=24, max_bin=64, rounds=3,
         params=dict(max_depth=5, eta=0.3, colsample_bytree=0.5)),
    dict(name="bylevel_05", kind="colsample", seed=342, n=20000, f=24, max_bin=64, rounds=3,
         params=dict(max_depth=6, eta=0.3, colsample_bylevel=0.5, seed=11)),
    dict(name="bylevel_06_bytree_07_subsample", kind="colsample", seed=343, n=20000, f=24, max_bin=64, rounds=3,
         params=dict(max_depth=5, eta=0.3, colsample_bylevel=0.6, colsample_bytree=0.7, subsample=0.8)),
    dict(name="gbs_03", kind="sampling", seed=321, n=20000, f=12, max_bin=64, rounds=3,
         params=dict(max_depth=5, eta=0.3, subsample=0.3, sampling_method="gradient_based")),
    dict(name="gbs_07_seed9", kind="sampling", seed=322, n=20000, f=12, max_bin=64, rounds=3,
         params=dict(max_depth=5, eta=0.3, subsample=0.7, sampling_method="gradient_based", seed=9)),
]


def bits(a):
    return [int(v) for v in np.asarray(a, dtype=np.float32).view(np.uint32)]


def run_case(case, kind):
    types = None
    if case["kind"] == "categorical":
        x, y, types = make_categorical_data(case)
    else:
        x, y = make_regression(case["n"], case["f"], seed=case["seed"])
    dm = rc.DMatrix.from_dense(x, label=y, kind=kind)
    if types is not None:
        dm.set_feature_types(types)
    params = {"tree_method": "hist", "device": "cuda:0" if kind == "cuda" else "cpu", "max_bin": case["max_bin"],
              "base_score": 0.0, "objective": "reg:squarederror", "lambda": 1.0, "min_child_weight": 1.0}
    params.update(case["params"])
    bst = rc.Booster(params, dm)
    for it in range(case["rounds"]):
        bst.update(it)
    cp, cv = dm.histogram_cuts()
    out = dict(case)
    out["feature_types"] = types
    out["cut_ptrs"] = [int(v) for v in cp]
    out["cut_values_bits"] = bits(cv)
    out["trees"] = []
    for t in bst.save_json()["learner"]["gradient_booster"]["model"]["trees"]:
        segs = dict(zip(t["categories_nodes"], zip(t["categories_segments"], t["categories_sizes"])))
        n_nodes = len(t["left_children"])
        out["trees"].append(dict(
            left=[int(v) for v in t["left_children"]], right=[int(v) for v in t["right_children"]],
            findex=[int(v) for v in t["split_indices"]], default_left=[int(v) for v in t["default_left"]],
            split_type=[int(v) for v in t["split_type"]], cond_bits=bits([float(v) for v in t["split_conditions"]]),
            base_weight_bits=bits([float(v) for v in t["base_weights"]]),
            loss_chg_bits=bits([float(v) for v in t["loss_changes"]]),
            sum_hess_bits=bits([float(v) for v in t["sum_hessian"]]),
            categories=[sorted(int(c) for c in t["categories"][segs[i][0]:segs[i][0] + segs[i][1]])
                        if i in segs else [] for i in range(n_nodes)]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/ref_gpu_next_trees.json")
    ap.add_argument("--kind", default="cuda", help="cpu = dry run of the plumbing on the CPU build (oracle/_ref)")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    res = {"_source": "reference gpu_hist (oracle/_ref_cuda/libxgboost.so = /root/reference built unmodified for "
                      "sm_100), run by tools/make_golden_gpu_next.py on a B200; inputs regenerate from the seeds via "
                      "tests/helpers.make_categorical_data / make_regression", "kind": args.kind, "cases": []}
    for c in CASES:
        res["cases"].append(run_case(c, args.kind))
        json.dump(res, open(args.out, "w"))  # keep what is done if the call is cut short
        print("case", c["name"], [len(t["left"]) for t in res["cases"][-1]["trees"]], flush=True)


if __name__ == "__main__":
    main()
