#!/usr/bin/env python
"""prediction on a quantised validation matrix -- the call GBTree::PredictBatchImpl makes for
every validation set each boosting iteration -- served by the reference's own gpu_predictor (EllpackLoader walk,
src/predictor/gpu_predictor.cu:180-351) or by our binding (xgboost_b200/csrc/shim/b200_predictor.cu ->
b200hist_forest_predict_reference_page, reading the reference's bit-packed page in place).

One child process per arm (the shim decides at load time, B200HIST_SHIM_PREDICTOR).  Prints one JSON line:
rows/s for margins and for leaf ids, and whether both arms returned the same bits.
    python tools/bench_predictor.py --rows 10000000 --features 64 --trees 50 --depth 6
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(args, out_path):
    import numpy as np
    import torch
    from oracle import ref_capi as rc
    rc.load("cuda")
    ctypes.CDLL(os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist.so"), mode=ctypes.RTLD_GLOBAL)
    shim = ctypes.CDLL(os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist_xgb_shim.so"), mode=ctypes.RTLD_GLOBAL)
    core = ctypes.CDLL(os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist.so"))
    core.b200hist_launch_count.restype = ctypes.c_uint64
    dev = torch.device("cuda", 0)
    f = args.features

    def batch(seed, m):
        g = torch.Generator(device=dev).manual_seed(seed)
        return torch.randn((m, f), generator=g, device=dev)
    w = torch.randn(f, generator=torch.Generator(device=dev).manual_seed(7), device=dev)
    xt = batch(1, args.train_rows)
    dtrain = rc.DMatrix.quantile_from_batches([xt], max_bin=256, kind="cuda")
    dtrain.set_label((xt @ w).cpu().numpy())
    nb = (args.rows + 999_999) // 1_000_000
    dvalid = rc.DMatrix.quantile_from_batches(
        rc.LazyBatches(nb, lambda b: batch(100 + b, min(1_000_000, args.rows - b * 1_000_000))), max_bin=256, kind="cuda")
    bst = rc.Booster({"tree_method": "hist", "device": "cuda:0", "max_bin": 256, "max_depth": args.depth, "eta": 0.1,
                      "objective": "reg:squarederror", "updater": "grow_gpu_hist_reference"}, dtrain)
    for i in range(args.trees):
        bst.update(i)
    torch.cuda.synchronize()
    res = {}
    before = int(core.b200hist_launch_count())
    for name, ptype in (("margin", 1), ("leaf", 6)):
        ts = []
        for rep in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = bst.predict_margin(dvalid, pred_type=ptype, iteration_begin=1 if ptype == 1 else 0)   # margins: no prediction cache; leaf ids are never cached
            ts.append(time.perf_counter() - t0)
        res[name] = {"s_best": min(ts[1:]), "s_all": ts, "rows_per_s": args.rows / min(ts[1:]),
                     "sha": __import__("hashlib").sha256(out.tobytes()).hexdigest()[:16], "n_out": int(out.size)}
    res["our_launches"] = int(core.b200hist_launch_count()) - before
    json.dump(res, open(out_path, "w"))
    del shim


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--train-rows", type=int, default=1_000_000)
    ap.add_argument("--features", type=int, default=64)
    ap.add_argument("--trees", type=int, default=50)
    ap.add_argument("--depth", type=int, default=6)
    ap.add_argument("--child", default=None)
    args = ap.parse_args()
    if args.child:
        child(args, args.child)
        return
    got = {}
    for gate in ("0", "1"):
        out = "/tmp/b200_predictor_%s.json" % gate
        env = dict(os.environ, B200HIST_SHIM_PREDICTOR=gate)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", out] +
                       [a for a in sys.argv[1:]], env=env, check=True, timeout=1800)
        got["ours" if gate == "1" else "reference"] = json.load(open(out))
    line = {"what": "PredictBatch / PredictLeaf on a QuantileDMatrix, %d x %d, %d trees of depth %d (includes the D2H of "
                    "the result inside XGBoosterPredictFromDMatrix, equal for both arms)" % (args.rows, args.features, args.trees, args.depth),
            "reference": got["reference"], "ours": got["ours"],
            "same_bits": {k: got["reference"][k]["sha"] == got["ours"][k]["sha"] for k in ("margin", "leaf")},
            "speedup": {k: got["reference"][k]["s_best"] / got["ours"][k]["s_best"] for k in ("margin", "leaf")}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
