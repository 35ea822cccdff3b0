#!/usr/bin/env python
"""Generate golden trees from the UNMODIFIED reference gpu_hist (oracle/_ref_cuda/libxgboost.so, built
from /root/reference by oracle/Makefile.ref_cuda) on a B200, and time it on the bench workloads.

Run on the GPU box:   python tools/make_golden_gpu.py --out gpurun_out/ref_gpu_trees.json [--time]
The resulting JSON is committed under tests/golden/; tests regenerate the inputs from the recorded
seeds (numpy default_rng) and compare trees split for split.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_capi as rc  # noqa: E402
from tests.helpers import SKETCH_CASES, make_regression, make_sparse_regression, make_sketch_batches  # noqa: E402

CASES = [
    dict(name="sq_d6_b64", n=20000, f=12, seed=101, max_bin=64, obj="sq", rounds=3,
         params=dict(max_depth=6, eta=0.3)),
    dict(name="sq_d8_b256", n=60000, f=32, seed=102, max_bin=256, obj="sq", rounds=2,
         params=dict(max_depth=8, eta=0.3)),
    dict(name="logistic_d5_b32", n=15000, f=9, seed=103, max_bin=32, obj="logistic", rounds=3,
         params=dict(max_depth=5, eta=0.5)),
    dict(name="sq_missing_d5", n=12000, f=10, seed=104, max_bin=32, obj="sq", rounds=2, missing=0.15,
         params=dict(max_depth=5, eta=0.3)),
    dict(name="sq_mcw_alpha", n=12000, f=10, seed=105, max_bin=32, obj="sq", rounds=2,
         params=dict(max_depth=5, eta=0.3, min_child_weight=40, alpha=0.5, **{"lambda": 2.0})),
    dict(name="sq_lossguide", n=12000, f=10, seed=106, max_bin=32, obj="sq", rounds=2,
         params=dict(max_depth=0, max_leaves=14, grow_policy="lossguide", eta=0.3)),
    dict(name="sq_wide_256f", n=30000, f=256, seed=107, max_bin=256, obj="sq", rounds=1,
         params=dict(max_depth=5, eta=0.3)),
    dict(name="sq_pair_128f_r3", n=25000, f=128, seed=109, max_bin=256, obj="sq", rounds=3,
         params=dict(max_depth=6, eta=0.5)),
    dict(name="sq_subsample07", n=20000, f=16, seed=110, max_bin=64, obj="sq", rounds=3,
         params=dict(max_depth=5, eta=0.3, subsample=0.7)),
    dict(name="sq_subsample05_seed9", n=20000, f=16, seed=111, max_bin=64, obj="sq", rounds=2,
         params=dict(max_depth=5, eta=0.3, subsample=0.5, seed=9)),
    dict(name="sparse_csr_d6", n=20000, f=300, seed=108, max_bin=64, obj="sq", rounds=2, sparse=0.03,
         params=dict(max_depth=6, eta=0.3)),
]


def gradients(obj, margin, y):
    if obj == "sq":
        g = (margin - y).astype(np.float32)
        h = np.ones_like(g)
    else:
        p = (1.0 / (1.0 + np.exp(-margin.astype(np.float64))))
        g = (p - y).astype(np.float32)
        h = np.maximum(p * (1.0 - p), 1e-16).astype(np.float32)
    return g, h


def run_case(case):
    import torch
    if case.get("sparse"):
        x, y, indptr, indices, values = make_sparse_regression(case["n"], case["f"], case["sparse"], case["seed"])
        dm = rc.DMatrix.from_csr(indptr, indices, values, case["f"], kind="cuda")
    else:
        x, y = make_regression(case["n"], case["f"], seed=case["seed"], missing=case.get("missing", 0.0))
        if case["obj"] == "logistic":
            y = (y > 0).astype(np.float32)
        dm = rc.DMatrix.from_cuda(torch.from_numpy(x).cuda(), kind="cuda")
    dm.set_label(y)
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": case["max_bin"], "base_score": 0.0,
              "objective": "reg:squarederror", "lambda": 1.0, "min_child_weight": 1.0}
    params.update(case["params"])
    bst = rc.Booster(params, dm)
    margin = np.zeros(case["n"], dtype=np.float32)
    margins = []
    for it in range(case["rounds"]):
        g, h = gradients(case["obj"], margin, y)
        bst.boost(it, torch.from_numpy(g).cuda(), torch.from_numpy(h).cuda())
        margin = bst.predict_margin(dm).astype(np.float32)
        margins.append(float(np.float64(margin).sum()))
    cp, cv = dm.histogram_cuts()
    trees = rc.trees_from_model(bst.save_json())
    # floats as uint32 bit patterns so the comparison is exact
    def bits(a):
        return [int(v) for v in np.asarray(a, dtype=np.float32).view(np.uint32)]
    out = dict(case)
    out["cut_ptrs"] = [int(v) for v in cp]
    out["cut_values_bits"] = bits(cv)
    out["margin_sums"] = margins
    out["trees"] = [dict(left=t["left"], right=t["right"], findex=t["findex"], default_left=t["default_left"],
                         cond_bits=bits(t["cond"]), base_weight_bits=bits(t["base_weight"]),
                         loss_chg_bits=bits(t["loss_chg"]), sum_hess_bits=bits(t["sum_hess"])) for t in trees]
    return out


def time_reference(rows, features, depth, max_bin=256, steps=5, warmup=2):
    """iters/s of the reference gpu_hist on the bench's synthetic workload (device batches -> QuantileDMatrix)."""
    import torch
    dev = torch.device("cuda:0")
    batch = 1_000_000
    nb = (rows + batch - 1) // batch
    w = torch.randn(features, generator=torch.Generator(device=dev).manual_seed(7), device=dev)

    def get(i):
        m = min(batch, rows - i * batch)
        g = torch.Generator(device=dev).manual_seed(2024 + i)
        return torch.randn((m, features), generator=g, device=dev)

    t0 = time.perf_counter()
    dm = rc.DMatrix.quantile_from_batches(rc.LazyBatches(nb, get), max_bin=max_bin, kind="cuda")
    y = np.empty(rows, dtype=np.float32)
    for i in range(nb):
        xb = get(i)
        y[i * batch:i * batch + xb.shape[0]] = (xb @ w).cpu().numpy()
    dm.set_label(y)
    t_build = time.perf_counter() - t0
    bst = rc.Booster({"tree_method": "hist", "device": "cuda:0", "max_depth": depth, "max_bin": max_bin,
                      "objective": "reg:squarederror", "eta": 0.3, "lambda": 1, "min_child_weight": 1}, dm)
    for i in range(warmup):
        bst.update(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        bst.update(warmup + i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return dict(rows=rows, features=features, depth=depth, iters_per_s=1.0 / dt, ms_per_iter=dt * 1e3,
                dmatrix_build_s=t_build)


def run_sketch_case(case):
    """Cuts of the reference's QuantileDMatrix over a stream of device batches (one sketch window per batch)."""
    import torch
    batches = [torch.from_numpy(b).cuda() for b in make_sketch_batches(case)]
    dm = rc.DMatrix.quantile_from_batches(batches, max_bin=case["max_bin"], kind="cuda")
    cp, cv = dm.histogram_cuts()
    out = {k: v for k, v in case.items() if k != "nan_cols"}
    if "nan_cols" in case:
        out["nan_cols"] = {str(k): v for k, v in case["nan_cols"].items()}
    out["cut_ptrs"] = [int(v) for v in cp]
    out["cut_values_bits"] = [int(v) for v in np.asarray(cv, dtype=np.float32).view(np.uint32)]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/ref_gpu_trees.json")
    ap.add_argument("--sketch-out", default=None,
                    help="write the QuantileDMatrix cut fixtures (tests/golden/ref_quantile_dmatrix_cuts.json) and exit")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--time-headline", action="store_true")
    args = ap.parse_args()
    if args.sketch_out:
        os.makedirs(os.path.dirname(args.sketch_out) or ".", exist_ok=True)
        res = {"_source": "cut values of the reference's QuantileDMatrix (device DataIter batches, "
                          "oracle/_ref_cuda/libxgboost.so = /root/reference built unmodified for sm_100) produced by "
                          "tools/make_golden_gpu.py --sketch-out on a B200; inputs regenerate from the seeds via "
                          "tests/helpers.make_sketch_batches", "cases": []}
        for c in SKETCH_CASES:
            res["cases"].append(run_sketch_case(c))
            print("sketch case", c["name"], res["cases"][-1]["cut_ptrs"][-1], "cuts", flush=True)
        json.dump(res, open(args.sketch_out, "w"))
        print("wrote", args.sketch_out)
        return
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    res = {"_source": "reference gpu_hist (oracle/_ref_cuda/libxgboost.so = /root/reference built unmodified for sm_100) "
                      "run by tools/make_golden_gpu.py on a B200; inputs regenerate from the seeds via "
                      "tests/helpers.make_regression", "cases": []}
    for c in CASES:
        res["cases"].append(run_case(c))
        print("case", c["name"], "trees", len(res["cases"][-1]["trees"]), "nodes",
              [len(t["left"]) for t in res["cases"][-1]["trees"]], flush=True)
    if args.time:
        res["timing"] = [time_reference(10_000_000, 64, 6)]
        print(res["timing"][-1], flush=True)
    if args.time_headline:
        res.setdefault("timing", []).append(time_reference(100_000_000, 256, 8, steps=3, warmup=1))
        print(res["timing"][-1], flush=True)
    json.dump(res, open(args.out, "w"))
    print("wrote", args.out)


if __name__ == "__main__":
    main()
