#!/usr/bin/env python
"""BASELINE.json config C1: demo/data/agaricus.txt.{train,test} (UCI mushrooms, 126 binary indicator features in libsvm
form; data, not source) -> tests/golden/agaricus.npz (CSR index arrays + labels, ~70 KB), plus the known answers of the
UNMODIFIED reference CPU library (oracle/_ref, tree_method=hist, binary:logistic, depth 6, 10 rounds) run here:
per-round train/test logloss and the model's trees.  BASELINE.md section 2 records 0.02390 / 0.02394 for the same run.

    python tools/make_golden_agaricus.py            (needs /root/reference and oracle/_ref/libxgboost.so)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def read_libsvm(path):
    indptr, indices, labels = [0], [], []
    for line in open(path):
        parts = line.split()
        labels.append(float(parts[0]))
        for tok in parts[1:]:
            k, v = tok.split(":")
            assert float(v) == 1.0
            indices.append(int(k))
        indptr.append(len(indices))
    return np.array(indptr, np.int32), np.array(indices, np.uint8), np.array(labels, np.float32)


def logloss(margin, y):
    p = 1.0 / (1.0 + np.exp(-margin.astype(np.float64)))
    p = np.clip(p, 1e-16, 1 - 1e-16)
    return float(-(y * np.log(p) + (1 - y) * np.log(1 - p)).mean())


def main():
    from oracle import ref_capi as rc
    d = "/root/reference/demo/data"
    tr = read_libsvm(os.path.join(d, "agaricus.txt.train"))
    te = read_libsvm(os.path.join(d, "agaricus.txt.test"))
    n_cols = 127
    out = os.path.join(ROOT, "tests", "golden", "agaricus.npz")
    np.savez_compressed(out, train_indptr=tr[0], train_indices=tr[1], train_label=tr[2], test_indptr=te[0],
                        test_indices=te[1], test_label=te[2], n_cols=np.int32(n_cols))

    def dm(part):
        ip, ix, y = part
        return rc.DMatrix.from_csr(ip, ix.astype(np.uint32), np.ones(ix.size, np.float32), n_cols, label=y, kind="cpu")
    dtr, dte = dm(tr), dm(te)
    bst = rc.Booster({"tree_method": "hist", "device": "cpu", "objective": "binary:logistic", "max_depth": 6,
                      "eta": 0.3, "nthread": 8}, dtr)   # everything else: the reference's defaults
    rounds = []
    for i in range(10):
        bst.update(i)
        rounds.append({"train_logloss": logloss(bst.predict_margin(dtr), tr[2]),
                       "test_logloss": logloss(bst.predict_margin(dte), te[2])})
    trees = bst.save_json()["learner"]["gradient_booster"]["model"]["trees"]
    keep = ("left_children", "right_children", "split_indices", "split_conditions", "default_left")
    json.dump({"generator": "tools/make_golden_agaricus.py", "reference": "oracle/_ref/libxgboost.so (CPU hist)",
               "params": {"objective": "binary:logistic", "max_depth": 6, "eta": 0.3, "rounds": 10},
               "rounds": rounds, "trees": [{k: t[k] for k in keep} for t in trees]},
              open(os.path.join(ROOT, "tests", "golden", "agaricus_known_answers.json"), "w"))
    print("train", tr[0].size - 1, "test", te[0].size - 1, "final", rounds[-1], "bytes", os.path.getsize(out))


if __name__ == "__main__":
    main()
