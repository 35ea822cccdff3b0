"""Property test of the whole updater against the oracle (the reference's own strategy for this path is property
based as well: tests/python-gpu/test_gpu_updaters.py:37-50 draws parameter sets with hypothesis).

Forty deterministic (derandomised) draws of data shape, symbol width, missing values, objective and TrainParam --
including row sampling, a tree-level feature set, per-depth feature sets and monotone constraints -- two boosting
rounds each; every node of every tree must equal the oracle's (split feature, threshold bits, default direction,
numbering, weights, hessian sums; loss_chg within the approximate-division tolerance) and so must the positions.
"""
import numpy as np
import pytest

import oracle as orc
from tests.helpers import logistic_gpair, make_cuts, make_regression, squared_error_gpair
from tests.test_gpu_parity import _compare_tree, _page


def _legitimate_tie(tree, o, cp, cv, p):
    """The oracle divides and multiplies in IEEE arithmetic, the device uses div.approx and fused multiply-adds
    (SURVEY.md 8(c)): where two different splits have gains that are equal in exact arithmetic -- e.g. every split
    whose two children are clipped to +-max_delta_step with gradients of one sign -- rounding picks the winner.  Returns
    True when the FIRST differing node is such a tie: the device's split scores within 8 float32 ulp of the oracle's
    best in the oracle's own gain table.  (Bit parity with the real reference is pinned by the golden and drop-in
    tests; this property test then stops following the diverged tree.)"""
    for i, (a, b) in enumerate(zip(tree.nodes, o["nodes"])):
        if b.is_leaf and a["is_leaf"]:
            continue
        same = (not b.is_leaf) and (not a["is_leaf"]) and a["findex"] == b.findex and \
            np.float32(a["fvalue"]).view(np.uint32) == np.float32(b.fvalue).view(np.uint32) and \
            a["default_left"] == bool(b.default_left)
        if same:
            continue
        if a["is_leaf"] or o["hist"] is None or not o["hist"][i].any():
            return False
        c = o["cands"][i]
        parent = (c.left_sum.g + c.right_sum.g, c.left_sum.h + c.right_sum.h)
        best, gains = orc.eval_node(o["hist"][i], parent, cp, cv, p, o["scale"], want_gains=True)
        f = a["findex"]
        if np.isneginf(a["fvalue"]):
            gbin = int(cp[f])
        else:
            hit = np.nonzero(cv[cp[f]:cp[f + 1]] == np.float32(a["fvalue"]))[0]
            if hit.size == 0:
                return False
            gbin = int(cp[f]) + int(hit[0]) + 1
        if gbin >= int(cp[f + 1]):
            return False
        g_dev = gains[gbin, 0 if a["default_left"] else 1]
        g_best = np.max(gains[np.isfinite(gains)])
        return abs(float(g_dev) - float(g_best)) <= 8 * np.finfo(np.float32).eps * max(abs(float(g_best)), 1e-30)
    return False

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

pytestmark = pytest.mark.gpu


@st.composite
def problems(draw):
    n = draw(st.sampled_from([1, 7, 33, 500, 2049, 6000]))
    f = draw(st.sampled_from([1, 3, 31, 32, 33, 64, 70]))
    max_bin = draw(st.sampled_from([2, 16, 64, 255, 256, 300]))
    missing = draw(st.sampled_from([0.0, 0.0, 0.15, 0.6]))
    obj = draw(st.sampled_from(["sq", "sq", "logistic"]))
    p = dict(eta=draw(st.sampled_from([0.1, 0.5, 1.0])), reg_lambda=draw(st.sampled_from([0.0, 1.0, 5.0])),
             reg_alpha=draw(st.sampled_from([0.0, 0.0, 0.4])), min_child_weight=draw(st.sampled_from([0.0, 1.0, 8.0])),
             min_split_loss=draw(st.sampled_from([0.0, 0.0, 0.5])))
    # max_delta_step is left out on purpose: with clipped weights EVERY split whose children have gradients of one
    # sign has the same gain in exact arithmetic, so IEEE (oracle) and fused/approximate (device) rounding pick
    # different winners all the time; that parameter is pinned against the real reference by the drop-in test and by
    # the dedicated oracle cases of test_gpu_parity.py
    if draw(st.booleans()):
        p.update(max_depth=draw(st.integers(1, 7)))
        if draw(st.booleans()):
            p.update(max_leaves=draw(st.integers(2, 40)))
    else:
        p.update(grow_policy="lossguide", max_leaves=draw(st.integers(2, 24)), max_depth=draw(st.sampled_from([0, 0, 5])))
    extra = dict(subsample=draw(st.sampled_from([1.0, 1.0, 0.6])), seed=draw(st.integers(0, 2**32 - 1)),
                 cols=draw(st.sampled_from(["all", "all", "tree", "level"])), mono=draw(st.sampled_from([False, False, True])),
                 data_seed=draw(st.integers(0, 10**6)))
    return n, f, max_bin, missing, obj, p, extra


@settings(max_examples=40, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture, HealthCheck.data_too_large])
@given(problems())
def test_random_problems_match_the_oracle(cuda_device, prob):
    import torch
    import xgboost_b200 as xb
    n, f, max_bin, missing, obj, p, extra = prob
    x, y = make_regression(n, f, seed=extra["data_seed"], missing=missing)
    yb = (y > 0).astype(np.float32)
    cp, cv = make_cuts(x, max_bin)
    page = _page(xb, x, cp, cv)
    gb = orc.bin_dense(x, cp, cv)
    up = xb.GPUHistMaker(page)
    up.configure(p)
    rng = np.random.default_rng(extra["data_seed"])
    kw = {}
    if extra["mono"]:
        kw["monotone"] = [int(v) for v in rng.integers(-1, 2, size=min(f, 5))]
        up.set_monotone_constraints(kw["monotone"])
    depth_rows = max(1, p.get("max_depth", 0) or 6)
    if extra["cols"] == "tree" and f > 1:
        kw["feature_set"] = np.sort(rng.choice(f, size=max(1, f // 2), replace=False)).astype(np.int32)
    if extra["cols"] == "level" and f > 1:
        kw["level_feature_sets"] = np.stack([np.sort(rng.choice(f, size=max(1, f // 2), replace=False))
                                             for _ in range(depth_rows)]).astype(np.int32)
    pred = np.zeros(n, dtype=np.float32)
    for it in range(2):
        seed = (extra["seed"] + it) & 0xFFFFFFFF
        if "level_feature_sets" in kw:
            up.set_tree_sampling(None, extra["subsample"], seed)
            up.set_level_feature_sets(kw["level_feature_sets"])
        else:
            up.set_tree_sampling(kw.get("feature_set"), extra["subsample"], seed)
        gp = squared_error_gpair(pred, y) if obj == "sq" else logistic_gpair(pred, yb)
        pos = torch.empty(n, dtype=torch.int32, device="cuda")
        tree = up.update(torch.from_numpy(gp).cuda(), pos)
        o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**p), want_hist=True, subsample=extra["subsample"],
                            sample_seed=seed, **kw)
        try:
            _compare_tree(xb, up, tree, o, cp, cv, hists=True)
        except AssertionError:
            # near-tie excuse only without per-node constraints / feature subsets (the gain table below has neither)
            if kw or not _legitimate_tie(tree, o, cp, cv, orc.Param.make(**p)):
                raise
            return
        assert np.array_equal(pos.cpu().numpy(), o["position"])
        pred = orc.update_prediction_cache(o["nodes"], o["position"], pred)
