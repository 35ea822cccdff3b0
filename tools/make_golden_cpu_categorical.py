#!/usr/bin/env python
"""Trees of the UNMODIFIED reference's CPU `hist` updater (oracle/_ref/libxgboost.so) on data with categorical
features -> tests/golden/ref_cpu_categorical_trees.json.  Runs in the build container (no GPU needed):

  python tools/make_golden_cpu_categorical.py --out tests/golden/ref_cpu_categorical_trees.json

Purpose: a structural pin of the oracle's categorical whole-tree restatement (oracle/hist_oracle.c
orc_update_tree_full) until the CUDA build of the reference can be run on such data.  The CPU updater sums float
histograms where gpu_hist sums fixed-point ones, so gains differ in the last bits; split features, category sets,
numerical split conditions (they are cut values) and the tree shape are comparable, default directions only where
the node really holds missing values of the split feature (otherwise both directions tie and the two updaters
break the tie differently).  Inputs regenerate from the seeds through make_data below.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_capi as rc  # noqa: E402
from tests.helpers import CATEGORICAL_N_CATS as N_CATS, make_categorical_data as make_data  # noqa: E402,F401

CASES = [dict(name="s%d_onehot%d_thr%d" % (s, oh, thr), seed=s, n=8000, max_bin=64, max_cat_to_onehot=oh,
              max_cat_threshold=thr, params=dict(max_depth=4, eta=0.5))
         for s, oh, thr in [(301, 4, 64), (301, 1, 64), (301, 100, 64), (302, 4, 64), (302, 100, 64), (303, 4, 8),
                            (304, 1, 3)]]


def bits(a):
    return [int(v) for v in np.asarray(a, dtype=np.float32).view(np.uint32)]


def run_case(case):
    x, y, types = make_data(case)
    dm = rc.DMatrix.from_dense(x, label=y, kind="cpu")
    dm.set_feature_types(types)
    params = {"tree_method": "hist", "device": "cpu", "max_bin": case["max_bin"], "base_score": 0.0,
              "objective": "reg:squarederror", "lambda": 1.0, "min_child_weight": 1.0,
              "max_cat_to_onehot": case["max_cat_to_onehot"], "max_cat_threshold": case["max_cat_threshold"]}
    params.update(case["params"])
    bst = rc.Booster(params, dm)
    bst.update(0)
    cp, cv = dm.histogram_cuts()
    t = bst.save_json()["learner"]["gradient_booster"]["model"]["trees"][0]
    segs = dict(zip(t["categories_nodes"], zip(t["categories_segments"], t["categories_sizes"])))
    n_nodes = len(t["left_children"])
    out = dict(case)
    out["feature_types"] = types
    out["cut_ptrs"] = [int(v) for v in cp]
    out["cut_values_bits"] = bits(cv)
    out["tree"] = dict(left=[int(v) for v in t["left_children"]], right=[int(v) for v in t["right_children"]],
                       findex=[int(v) for v in t["split_indices"]], default_left=[int(v) for v in t["default_left"]],
                       split_type=[int(v) for v in t["split_type"]],
                       cond_bits=bits([float(v) for v in t["split_conditions"]]),
                       categories=[sorted(int(c) for c in t["categories"][segs[i][0]:segs[i][0] + segs[i][1]])
                                   if i in segs else [] for i in range(n_nodes)])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="tests/golden/ref_cpu_categorical_trees.json")
    args = ap.parse_args()
    res = {"_source": "reference CPU hist (oracle/_ref/libxgboost.so = /root/reference built unmodified) on categorical "
                      "data, by tools/make_golden_cpu_categorical.py; inputs regenerate from the seeds via make_data",
           "cases": [run_case(c) for c in CASES]}
    json.dump(res, open(args.out, "w"))
    for c in res["cases"]:
        print(c["name"], len(c["tree"]["left"]), "nodes,", sum(c["tree"]["split_type"]), "categorical splits")


if __name__ == "__main__":
    main()
