"""GPU parity tests of code written AFTER this round's GPU minutes were spent: compiled, reviewed, never run on
hardware.  They are skipped unless B200HIST_RUN_UNVALIDATED=1 so that the suite that has been seen green on a B200
stays the suite that runs by default; the first GPU call of the next round runs them
(`B200HIST_RUN_UNVALIDATED=1 python -m pytest tests/test_gpu_next.py -m gpu -x -q`) and a test moves to
test_gpu_parity.py / test_reference_golden.py once it has passed there.

Their checkers are already pinned on the CPU: the oracle functions they compare with reproduce the reference's
literal unit-test expectations and, for whole trees, trees of the unmodified reference gpu_hist
(tests/test_oracle_golden.py, tests/test_reference_golden.py).
"""
import json
import os

import numpy as np
import pytest

import oracle as orc
from tests.helpers import make_cuts, make_regression, squared_error_gpair
from tests.test_gpu_parity import _compare_tree, _page, xb  # noqa: F401  (xb is a fixture)

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("B200HIST_RUN_UNVALIDATED") != "1",
                                 reason="not yet run on hardware; set B200HIST_RUN_UNVALIDATED=1")]

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---------------------------------------------------------------------------------------------------------------------
# Feature interaction constraints (b200hist_updater_set_interaction_constraints, csrc/updater.cu IcLevelKernel)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", json.load(open(os.path.join(GOLDEN, "ref_gpu_interaction_trees.json")))["cases"],
                         ids=lambda c: c["name"])
def test_interaction_constraints_match_reference_gpu_hist_bit_for_bit(xb, case):
    """Overlapping constraint sets: the reference's trees depend on its shared Query buffer (include/b200hist.h);
    ours must be the same trees, bit for bit, loss_chg included."""
    import torch
    from tools.probe_interaction_reference import make_data
    x, y = make_data(case)
    cp = np.array(case["cut_ptrs"], dtype=np.uint32)
    cv = np.array(case["cut_values_bits"], dtype=np.uint32).view(np.float32)
    page = _page(xb, x, cp, cv)
    up = xb.GPUHistMaker(page)
    up.configure(dict(case["params"]))
    up.set_interaction_constraints(case["sets"])
    margin = torch.zeros(case["n"], dtype=torch.float32, device="cuda")
    yd = torch.from_numpy(y).cuda()
    for t, ref in enumerate(case["trees"]):
        gp = torch.stack([margin - yd, torch.ones_like(yd)], dim=1).contiguous()
        tree = up.update(gp)
        assert tree.num_nodes == len(ref["left"]), "tree %d node count" % t
        for i, nd in enumerate(tree.nodes):
            assert (nd["left"], nd["right"]) == (ref["left"][i], ref["right"][i]), (t, i)
            if not nd["is_leaf"]:
                assert nd["findex"] == ref["findex"][i], (t, i)
                assert np.float32(nd["fvalue"]).view(np.uint32) == ref["cond_bits"][i], (t, i)
                assert int(nd["default_left"]) == ref["default_left"][i], (t, i)
                assert np.float32(nd["loss_chg"]).view(np.uint32) == ref["loss_chg_bits"][i], (t, i)
            else:
                assert np.float32(nd["leaf_value"]).view(np.uint32) == ref["cond_bits"][i], (t, i)
            assert np.float32(nd["sum_hess"]).view(np.uint32) == ref["sum_hess_bits"][i], (t, i)
            assert np.float32(nd["base_weight"]).view(np.uint32) == ref["base_weight_bits"][i], (t, i)
        up.update_prediction_cache(margin)


IC_CASES = [
    dict(name="disjoint_d5", f=9, sets=[[0, 1, 2], [3, 4, 5], [6, 7]], params=dict(max_depth=5, eta=0.5)),
    dict(name="chain_d6", f=8, sets=[[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [5, 6], [6, 7]],
         params=dict(max_depth=6, eta=0.5)),
    dict(name="chain_lossguide", f=8, sets=[[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [5, 6], [6, 7]],
         params=dict(max_depth=0, max_leaves=24, grow_policy="lossguide", eta=0.5)),
    dict(name="wide_130f", f=130, sets=[[0, 1, 129], [129, 64, 65], [65, 100], [2, 3]],
         params=dict(max_depth=5, eta=0.5)),                                   # more than one / two 64-bit words
    dict(name="free_features", f=6, sets=[[0, 1]], params=dict(max_depth=4, eta=0.5)),
    dict(name="chain_bylevel", f=8, sets=[[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [5, 6], [6, 7]], level_sets=True,
         params=dict(max_depth=5, eta=0.5)),
    dict(name="chain_monotone", f=8, sets=[[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [5, 6], [6, 7]],
         mono=[1, -1, 0, 1, 0, -1, 0, 0], params=dict(max_depth=5, eta=0.5)),
]


@pytest.mark.parametrize("case", IC_CASES, ids=lambda c: c["name"])
def test_interaction_constraints_match_oracle(xb, case):
    import torch
    n, f = 12000, case["f"]
    x, y = make_regression(n, f, seed=61)
    rng = np.random.default_rng(62)
    y = (y + x[:, 0] * x[:, 1 % f] - x[:, 2 % f] * x[:, 3 % f] + x[:, (f - 1)] * x[:, (f - 2)] +
         0.1 * rng.normal(size=n)).astype(np.float32)
    cp, cv = make_cuts(x, 32)
    page = _page(xb, x, cp, cv)
    gb = orc.bin_dense(x, cp, cv)
    params = dict(case["params"])
    up = xb.GPUHistMaker(page)
    up.configure(params)
    up.set_interaction_constraints(case["sets"])
    level_sets = None
    if case.get("level_sets"):  # colsample_bylevel draws: 5 of 8 features per depth, ascending or not
        level_sets = np.stack([rng.permutation(f)[:5] for _ in range(params["max_depth"] + 1)]).astype(np.int32)
        up.set_level_feature_sets(level_sets)
    if case.get("mono"):
        up.set_monotone_constraints(case["mono"])
    pred = np.zeros(n, dtype=np.float32)
    for it in range(3):
        gp = squared_error_gpair(pred, y)
        tree = up.update(torch.from_numpy(gp).cuda())
        o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), interaction=case["sets"], want_hist=True,
                            level_feature_sets=level_sets, monotone=case.get("mono"))
        _compare_tree(xb, up, tree, o, cp, cv, hists=True)
        pred = orc.update_prediction_cache(o["nodes"], o["position"], pred)
    # removing the constraints restores the plain tree
    up.set_interaction_constraints(None)
    if level_sets is not None:
        up.set_level_feature_sets(None)
    if case.get("mono"):
        up.set_monotone_constraints(None)
    gp = squared_error_gpair(np.zeros(n, np.float32), y)
    t_plain = up.update(torch.from_numpy(gp).cuda())
    o_plain = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params))
    _compare_tree(xb, up, t_plain, o_plain, cp, cv, hists=False)
