"""Outputs RECORDED on a B200 for the paths finished after the round's regular GPU sessions, checked on the CPU.

tests/golden/b200_quick_check/ holds what tools/quick_check/quick_gpu_check (C ABI only) wrote in the last GPU
seconds of round 1 (profiles/r01_quick_check_gpurun.json): node arrays, row positions and category sets of five
trees -- plain, overlapping interaction constraints, plain again after removing them, categorical features with
sorted-partition splits, categorical features forced one-hot.  The inputs regenerate from seeds
(tools/quick_check/make_inputs.py).  Every tree must equal the oracle's: structure, split conditions and base
weights bit for bit, category sets, and the position of every row (i.e. the partitioner and FinalisePosition
routed by category set correctly).  This is evidence about the code as it was when recorded; the live GPU tests
for these paths are tests/test_gpu_next.py.
"""
import os

import numpy as np

from tests import quick_check_compare as compare

REC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "b200_quick_check")


def test_recorded_b200_trees_equal_the_oracle():
    res = compare.run(REC)
    assert res == {k: True for k in ("plain", "interaction", "plain_again", "categorical", "categorical_onehot")}, res


def test_recorded_interaction_tree_follows_the_shared_query_buffer():
    """The recording discriminates: it is NOT the tree of the per-query reading, and not the unconstrained tree."""
    import oracle as orc
    num, _ = compare.inputs()
    p = orc.Param.make(max_depth=5, eta=0.5)
    gb = orc.bin_dense(num["x"], num["cp"], num["cv"])
    got = [int(v) for v in np.fromfile(os.path.join(REC, "interaction_nodes.bin"), dtype=compare.NODE)["findex"]]
    chain = [[0, 1], [1, 2], [2, 3], [3, 4], [4, 5]]
    per_query = orc.update_tree(gb, num["cp"], num["cv"], num["gpair"], p, interaction=chain,
                                interaction_shared_buffer=False)
    plain = [int(v) for v in np.fromfile(os.path.join(REC, "plain_nodes.bin"), dtype=compare.NODE)["findex"]]
    assert got != [nd.findex for nd in per_query["nodes"]] and got != plain


def test_recorded_categorical_trees_hold_categorical_splits():
    for name in ("categorical", "categorical_onehot"):
        assert int(np.fromfile(os.path.join(REC, name + "_is_cat.bin"), dtype=np.int32).sum()) >= 10
