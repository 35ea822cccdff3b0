#!/usr/bin/env python
"""One short GPU call: trees of the UNMODIFIED reference gpu_hist (oracle/_ref_cuda) under OVERLAPPING interaction
constraints, to settle which reading of FeatureInteractionConstraintDevice::Query its trees follow (DESIGN (f) 4,
oracle orc_update_tree_ic shared_result_buffer 0/1).  No torch import (host DMatrix, device training).

  python tools/probe_interaction_reference.py --out gpurun_out/ref_interaction_trees.json
The output is committed as tests/golden/ref_gpu_interaction_trees.json; inputs regenerate from the seeds below.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_capi as rc  # noqa: E402
from tests.helpers import make_interaction_data as make_data  # noqa: E402

SETS = [[0, 1], [1, 2], [2, 3], [3, 4], [4, 5]]
CASES = [dict(name="chain_%d" % s, seed=s, n=6000, f=6, max_bin=32, rounds=2,
              params=dict(max_depth=5, eta=0.5)) for s in (201, 202, 203)]


def bits(a):
    return [int(v) for v in np.asarray(a, dtype=np.float32).view(np.uint32)]


def run_case(case, kind="cuda"):
    x, y = make_data(case)
    dm = rc.DMatrix.from_dense(x, label=y, kind=kind)
    params = {"tree_method": "hist", "device": "cuda:0" if kind == "cuda" else "cpu", "max_bin": case["max_bin"], "base_score": 0.0,
              "objective": "reg:squarederror", "lambda": 1.0, "min_child_weight": 1.0,
              "interaction_constraints": json.dumps(SETS)}
    params.update(case["params"])
    bst = rc.Booster(params, dm)
    for it in range(case["rounds"]):
        bst.update(it)
    cp, cv = dm.histogram_cuts()
    trees = rc.trees_from_model(bst.save_json())
    out = dict(case)
    out["sets"] = SETS
    out["cut_ptrs"] = [int(v) for v in cp]
    out["cut_values_bits"] = bits(cv)
    out["trees"] = [dict(left=t["left"], right=t["right"], findex=t["findex"], default_left=t["default_left"],
                         cond_bits=bits(t["cond"]), base_weight_bits=bits(t["base_weight"]),
                         loss_chg_bits=bits(t["loss_chg"]), sum_hess_bits=bits(t["sum_hess"])) for t in trees]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/ref_interaction_trees.json")
    ap.add_argument("--kind", default="cuda", help="cpu = dry run of the plumbing on the CPU build (oracle/_ref)")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    res = {"_source": "reference gpu_hist (oracle/_ref_cuda/libxgboost.so = /root/reference built unmodified for "
                      "sm_100) with interaction_constraints=%s, run by tools/probe_interaction_reference.py on a "
                      "B200; inputs regenerate from the seeds via make_data in that script" % json.dumps(SETS),
           "cases": []}
    for c in CASES:
        res["cases"].append(run_case(c, args.kind))
        json.dump(res, open(args.out, "w"))      # keep what is done if the call is cut short
        print("case", c["name"], [len(t["left"]) for t in res["cases"][-1]["trees"]], flush=True)


if __name__ == "__main__":
    main()
