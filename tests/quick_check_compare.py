#!/usr/bin/env python
"""Checks the *.bin files tools/quick_check/quick_gpu_check wrote on a B200 against the oracle (test infrastructure:
the oracle is only ever imported from tests/).
  python tests/quick_check_compare.py [directory]     (default gpurun_out/quick_check; the run kept from round 1 is
                                                       tests/golden/b200_quick_check)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402
from tests.test_reference_golden import categorical_global_bins  # noqa: E402
from tools.quick_check.make_inputs import inputs  # noqa: E402

NODE = np.dtype([("parent", "<i4"), ("left", "<i4"), ("right", "<i4"), ("findex", "<i4"), ("fvalue", "<f4"),
                 ("default_left", "<i4"), ("is_leaf", "<i4"), ("leaf_value", "<f4"), ("loss_chg", "<f4"),
                 ("sum_hess", "<f4"), ("base_weight", "<f4"), ("depth", "<i4")])
OUT = os.path.join(ROOT, "gpurun_out", "quick_check")


def check(name, o, cats=False, out=None):
    out = out or OUT
    nodes = np.fromfile(os.path.join(out, name + "_nodes.bin"), dtype=NODE)
    pos = np.fromfile(os.path.join(out, name + "_pos.bin"), dtype=np.int32)
    bad = []
    if len(nodes) != len(o["nodes"]):
        bad.append("node count %d vs %d" % (len(nodes), len(o["nodes"])))
    else:
        is_cat = np.fromfile(os.path.join(out, name + "_is_cat.bin"), dtype=np.int32) if cats else None
        bits = np.fromfile(os.path.join(out, name + "_cat_bits.bin"), dtype=np.uint32).reshape(len(nodes), -1) if cats else None
        for i, (a, b) in enumerate(zip(nodes, o["nodes"])):
            ok = (a["left"], a["right"], bool(a["is_leaf"])) == (b.left, b.right, bool(b.is_leaf))
            if ok and not b.is_leaf:
                ok = a["findex"] == b.findex and bool(a["default_left"]) == bool(b.default_left)
                if cats and o["is_cat"][i]:
                    got = {32 * w + (31 - k) for w in range(bits.shape[1]) for k in range(32) if (int(bits[i, w]) >> k) & 1}
                    ok = ok and is_cat[i] == 1 and np.isnan(a["fvalue"]) and got == o["cat_sets"][i]
                else:
                    ok = ok and np.float32(a["fvalue"]).view(np.uint32) == np.float32(b.fvalue).view(np.uint32)
            ok = ok and np.float32(a["base_weight"]).view(np.uint32) == np.float32(b.base_weight).view(np.uint32)
            if not ok:
                bad.append("node %d" % i)
        if not np.array_equal(pos, o["position"]):
            bad.append("positions")
    print(name, "OK" if not bad else "MISMATCH " + ", ".join(bad[:6]), "(%d nodes)" % len(nodes))
    return not bad


def run(out=None):
    """All five cases; returns {case: bool}."""
    num, cat = inputs()
    p = orc.Param.make(max_depth=5, eta=0.5)
    gb = orc.bin_dense(num["x"], num["cp"], num["cv"])
    plain = orc.update_tree(gb, num["cp"], num["cv"], num["gpair"], p)
    chain = [[0, 1], [1, 2], [2, 3], [3, 4], [4, 5]]
    res = {"plain": check("plain", plain, out=out)}
    res["interaction"] = check("interaction", orc.update_tree(gb, num["cp"], num["cv"], num["gpair"], p,
                                                              interaction=chain), out=out)
    res["plain_again"] = check("plain_again", plain, out=out)
    flags = [int(v) for v in cat["is_cat"]]
    gbc = categorical_global_bins(cat["x"], cat["cp"], cat["cv"], flags)
    res["categorical"] = check("categorical", orc.update_tree(gbc, cat["cp"], cat["cv"], cat["gpair"], p,
                                                              feature_is_cat=flags), True, out=out)
    res["categorical_onehot"] = check("categorical_onehot",
                                      orc.update_tree(gbc, cat["cp"], cat["cv"], cat["gpair"], p, feature_is_cat=flags,
                                                      max_cat_to_onehot=100), True, out=out)
    return res


if __name__ == "__main__":
    sys.exit(0 if all(run(sys.argv[1] if len(sys.argv) > 1 else None).values()) else 1)
