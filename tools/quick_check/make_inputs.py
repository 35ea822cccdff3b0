#!/usr/bin/env python
"""Inputs of tools/quick_check/quick_gpu_check.cu (numerical data with an interaction chain, categorical data of
the CPU golden fixture's recipe) as raw binary files under tools/quick_check/inputs/."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.helpers import make_categorical_data as cat_data  # noqa: E402
from tests.helpers import make_cuts  # noqa: E402
from tests.helpers import make_interaction_data as num_data  # noqa: E402

OUT = os.path.join(ROOT, "tools", "quick_check", "inputs")


def dump(name, x, gpair, cp, cv, has_missing):
    np.array([x.shape[0], x.shape[1], 1 if has_missing else 0], dtype=np.uint64).tofile(os.path.join(OUT, name + "_meta.bin"))
    np.ascontiguousarray(x, dtype=np.float32).tofile(os.path.join(OUT, name + "_x.bin"))
    np.ascontiguousarray(gpair, dtype=np.float32).tofile(os.path.join(OUT, name + "_gpair.bin"))
    np.ascontiguousarray(cp, dtype=np.uint32).tofile(os.path.join(OUT, name + "_cut_ptrs.bin"))
    np.ascontiguousarray(cv, dtype=np.float32).tofile(os.path.join(OUT, name + "_cut_values.bin"))


def inputs():
    x, y = num_data(dict(seed=201, n=6000, f=6))
    cp, cv = make_cuts(x, 32)
    num = dict(x=x, gpair=np.stack([-y, np.ones_like(y)], 1).astype(np.float32), cp=cp, cv=cv)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_cpu_categorical_trees.json")))["cases"][0]
    xc, yc, types = cat_data(gold)
    is_cat = np.array([1 if t == "c" else 0 for t in types], dtype=np.uint8)
    cat = dict(x=xc, gpair=np.stack([-yc, np.ones_like(yc)], 1).astype(np.float32),
               cp=np.array(gold["cut_ptrs"], dtype=np.uint32),
               cv=np.array(gold["cut_values_bits"], dtype=np.uint32).view(np.float32), is_cat=is_cat)
    return num, cat


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    num, cat = inputs()
    dump("num", num["x"], num["gpair"], num["cp"], num["cv"], False)
    xs = cat["x"].copy()
    xs[:, cat["is_cat"] != 0] -= 0.5      # numerical binning then puts category c into bin c
    dump("cat", xs, cat["gpair"], cat["cp"], cat["cv"], True)
    cat["is_cat"].tofile(os.path.join(OUT, "cat_is_cat.bin"))
    print("wrote", OUT)
