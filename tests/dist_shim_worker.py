"""One worker of the 2-process drop-in test (tests/test_multi_gpu.py): a job of the UNMODIFIED reference library with the
reference's own tracker and communicator, data split by rows, one GPU per worker.  The updater is either the reference's
("grow_gpu_hist_reference") or ours through the shim ("grow_b200_hist"); worker 0 writes the model.

usage: python -m tests.dist_shim_worker <task_id> <tracker_args_json> <updater> <out.json> <extra_params_json>"""
import ctypes
import json
import os
import sys

import numpy as np


def main(task_id: int, tracker_args: dict, updater: str, out_path: str, extra: dict) -> None:
    import torch
    from oracle import ref_capi as rc
    from tests.conftest import ROOT
    from tests.helpers import make_regression
    rc.load("cuda")
    ctypes.CDLL(os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist.so"), mode=ctypes.RTLD_GLOBAL)
    shim = ctypes.CDLL(os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist_xgb_shim.so"), mode=ctypes.RTLD_GLOBAL)
    rc.communicator_init("cuda", dict(tracker_args, dmlc_task_id=str(task_id), dmlc_communicator="rabit"))
    try:
        rank = rc.communicator_rank("cuda")
        world = 2
        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        n = int(extra.pop("rows", 60000))
        x, y = make_regression(n, 24, seed=91, missing=float(extra.pop("missing", 0.03)))
        lo, hi = (n * rank) // world + (17 if rank else 0), (n * (rank + 1)) // world + (17 if rank + 1 < world else 0)
        if extra.get("objective") == "binary:logistic":
            y = (y > 0).astype(np.float32)
        xs, ys = x[lo:hi], y[lo:hi]                      # uneven shards
        dm = rc.DMatrix.from_cuda(torch.from_numpy(xs).to("cuda:%d" % dev), kind="cuda")
        dm.set_label(ys)
        params = {"tree_method": "hist", "device": "cuda:%d" % dev, "max_bin": 64, "max_depth": 6, "eta": 0.3,
                  "objective": "reg:squarederror", "base_score": 0.5, "updater": updater}
        params.update(extra)
        bst = rc.Booster(params, dm)
        for i in range(4):
            bst.update(i)
        trees = bst.save_json()["learner"]["gradient_booster"]["model"]["trees"]
        keep = ("left_children", "right_children", "split_indices", "split_conditions", "default_left", "base_weights",
                "loss_changes", "sum_hessian")
        margin = bst.predict_margin(dm)
        lib = ctypes.CDLL(os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist.so"))
        lib.b200hist_launch_count.restype = ctypes.c_uint64
        json.dump({"trees": [{k: t[k] for k in keep} for t in trees], "rank": rank, "rows": [lo, hi],
                   "margin_bits": margin.view(np.uint32).tolist()[:2000], "launches": int(lib.b200hist_launch_count())},
                  open("%s.rank%d" % (out_path, rank), "w"))
        del bst, dm
    finally:
        rc.communicator_finalize("cuda")
    del shim


if __name__ == "__main__":
    main(int(sys.argv[1]), json.loads(sys.argv[2]), sys.argv[3], sys.argv[4], json.loads(sys.argv[5]))
