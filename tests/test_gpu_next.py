"""GPU parity tests of the updater's remaining parameters -- interaction constraints, categorical splits, colsample_bynode,
gradient-based sampling, prediction on quantised pages, several trees per iteration -- and their drop-in tests through
the shim.  Written at the end of round 1 behind a gate; first run on a B200 at the start of round 2
(gpurun_out/next_tests.log of that call: 76 of 84 green, the other 8 were two test defects, one count assumption and
the multi-class prediction-cache stride), ungated since.

Their checkers are pinned on the CPU: the oracle functions they compare with reproduce the reference's
literal unit-test expectations and, for whole trees, trees of the unmodified reference gpu_hist
(tests/test_oracle_golden.py, tests/test_reference_golden.py).
"""
import json
import os

import numpy as np
import pytest

import oracle as orc
from tests.helpers import make_cuts, make_regression, squared_error_gpair
from tests.test_gpu_parity import _compare_tree, _global_bins, _page, xb  # noqa: F401  (xb is a fixture)

pytestmark = pytest.mark.gpu

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
    from tests.helpers import make_interaction_data as make_data
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
    if case.get("level_sets"):  # colsample_bylevel draws: 5 of 8 features per depth
        # ascending, like every set the reference's ColumnSampler hands out (src/common/random.cu:98 thrust::sort,
        # random.cc:71 std::sort); the ABI refuses anything else
        level_sets = np.stack([np.sort(rng.permutation(f)[:5]) for _ in range(params["max_depth"] + 1)]).astype(np.int32)
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


# ---------------------------------------------------------------------------------------------------------------------
# Categorical split evaluation (b200hist_evaluate_splits_cat: CatFeatureScanKernel / CatFinishKernel, evaluate.cu)
# ---------------------------------------------------------------------------------------------------------------------
def _cats_of(bits_row):
    return {32 * w + (31 - b) for w in range(bits_row.size) for b in range(32) if (int(bits_row[w]) >> b) & 1}


def test_categorical_evaluation_known_answers(xb):
    """The reference's own literal vectors (tests/cpp/tree/gpu_hist/test_evaluate_splits.cu:39-335), the same twelve
    the oracle is pinned on."""
    import torch
    from tests.test_oracle_golden import _dummy_scale, _to_fixed
    spec = json.load(open(os.path.join(GOLDEN, "eval_categorical_known_answers.json")))
    scale = _dummy_scale()
    q = xb.lib.Quantiser(*[float(v) for v in scale])
    for case in spec["cases"]:
        name = case["name"]
        cp = np.array(case["cut_ptrs"], dtype=np.uint32)
        cv = np.array(case["cut_values"], dtype=np.float32)
        page = xb.EllpackPage(xb.HistogramCuts(cp, cv), 1, 0, False)
        hist = _to_fixed(case["hist"], scale)
        parent = _to_fixed([case["parent"]], scale, as_float32=False)[0]
        p = xb.TrainParam.make(**case["param"])
        cands, is_cat, bits = xb.components.evaluate_splits_cat(
            page, torch.from_numpy(hist).cuda()[None].contiguous(), [tuple(int(v) for v in parent)], p, q,
            case["is_cat"], max_cat_to_onehot=case["max_cat_to_onehot"], n_cat_words=max(1, int(cv.max()) // 32 + 1))
        c, cats = cands[0], _cats_of(bits[0])
        o, o_is_cat, _, o_cats = orc.eval_node_cat(hist, parent, cp, cv, orc.Param.make(**case["param"]), scale,
                                                   case["is_cat"], case["max_cat_to_onehot"])
        e = case["expect"]
        if "findex" in e:
            assert c.findex == e["findex"], name
        if "dir" in e:
            assert c.dir == e["dir"], name
        if "cats" in e:
            assert cats == set(e["cats"]), (name, cats)
        if e.get("is_cat"):
            assert is_cat[0] == 1, name
        if e.get("fvalue_is_nan"):
            assert np.isnan(c.fvalue), name
        if e.get("sum_is_parent"):
            assert (c.left_sum.grad + c.right_sum.grad, c.left_sum.hess + c.right_sum.hess) == tuple(parent), name
        if "loss_chg" in e:
            assert abs(c.loss_chg - e["loss_chg"]) <= e.get("loss_chg_tol", 1e-6), (name, c.loss_chg)
        # and everything the oracle reports
        assert (c.findex, c.dir, bool(is_cat[0]), cats) == (o.findex, o.dir, o_is_cat, o_cats), name
        assert (c.left_sum.grad, c.left_sum.hess, c.right_sum.grad, c.right_sum.hess) == \
            (o.left_sum.g, o.left_sum.h, o.right_sum.g, o.right_sum.h), name


@pytest.mark.parametrize("onehot,threshold,params", [
    (4, 64, dict()), (1, 64, dict(reg_lambda=0.0, min_child_weight=0.0)), (100, 64, dict(reg_alpha=0.3)),
    (4, 5, dict(min_child_weight=20.0)), (4, 64, dict(max_delta_step=0.7)), (40, 33, dict(reg_lambda=3.0))])
def test_categorical_evaluation_random_histograms(xb, onehot, threshold, params):
    """Mixed numerical / categorical features with 2..70 categories (one-hot and sorted partition, more than one
    32-bin chunk), several nodes of different sizes, missing values; candidate, flag and category set against the
    oracle.  Gains are compared with the tolerance of the numerical test (approximate vs IEEE division)."""
    import torch
    rng = np.random.default_rng(17)
    n = 8000
    n_cats = [2, 3, 5, 9, 31, 33, 70]
    cols, cut_ptrs, cut_values, is_cat = [], [0], [], []
    for k in n_cats:  # categorical columns: codes 0..k-1, 8 % missing
        col = rng.integers(0, k, size=n).astype(np.float32)
        col[rng.random(n) < 0.08] = np.nan
        cols.append(col)
        cut_values += [float(v) for v in range(k)]
        cut_ptrs.append(len(cut_values))
        is_cat.append(1)
    xnum, _ = make_regression(n, 4, seed=23, missing=0.05)
    ncp, ncv = make_cuts(xnum, 32)
    for j in range(4):
        cols.append(xnum[:, j])
        cut_values += [float(v) for v in ncv[ncp[j]:ncp[j + 1]]]
        cut_ptrs.append(len(cut_values))
        is_cat.append(0)
    cp = np.array(cut_ptrs, dtype=np.uint32)
    cv = np.array(cut_values, dtype=np.float32)
    # global bins: a category is its own bin (SearchCatBin), numerical features through the oracle's SearchBin
    gb = np.full((n, len(cols)), -1, dtype=np.int32)
    for j, col in enumerate(cols):
        ok = ~np.isnan(col)
        if is_cat[j]:
            gb[ok, j] = cp[j] + col[ok].astype(np.int32)
        else:
            gb[ok, j] = orc.bin_dense(col[:, None], np.array([0, cp[j + 1] - cp[j]], dtype=np.uint32),
                                      cv[cp[j]:cp[j + 1]])[ok, 0] + cp[j]
    effect = rng.normal(size=70)
    y = (effect[np.nan_to_num(cols[6], nan=0).astype(int)] + 0.5 * effect[np.nan_to_num(cols[4], nan=0).astype(int)] +
         np.nan_to_num(xnum[:, 0]) + 0.3 * rng.normal(size=n)).astype(np.float32)
    gp = squared_error_gpair(np.zeros(n, np.float32), y)
    scale, _ = orc.quantiser(gp)
    gq = orc.quantise(gp, scale)
    subsets = [np.arange(n), rng.choice(n, 900, replace=False), rng.choice(n, 60, replace=False)]
    hists = np.stack([orc.build_hist(gb, gq, s.astype(np.uint32), cv.size) for s in subsets])
    parents = [tuple(int(v) for v in gq[s].sum(axis=0)) for s in subsets]
    q = xb.lib.Quantiser(*[float(v) for v in scale])
    page = xb.EllpackPage(xb.HistogramCuts(cp, cv), 1, 0, False)
    cands, flags, bits = xb.components.evaluate_splits_cat(
        page, torch.from_numpy(hists).cuda(), parents, xb.TrainParam.make(**params), q, is_cat,
        max_cat_to_onehot=onehot, max_cat_threshold=threshold, n_cat_words=3)
    op = orc.Param.make(**params)
    for k, c in enumerate(cands):
        o, o_is_cat, _, o_cats = orc.eval_node_cat(hists[k], parents[k], cp, cv, op, scale, is_cat, onehot, threshold)
        name = "node %d onehot %d threshold %d %r" % (k, onehot, threshold, params)
        assert (c.findex, c.dir, bool(flags[k])) == (o.findex, o.dir, o_is_cat), name
        assert (c.left_sum.grad, c.left_sum.hess, c.right_sum.grad, c.right_sum.hess) == \
            (o.left_sum.g, o.left_sum.h, o.right_sum.g, o.right_sum.h), name
        if o_is_cat:
            assert _cats_of(bits[k]) == o_cats, name
            assert np.isnan(c.fvalue) and c.split_gbin == -1, name
        else:
            assert np.float32(c.fvalue).view(np.uint32) == np.float32(o.fvalue).view(np.uint32), name
        assert np.isclose(c.loss_chg, o.loss_chg, rtol=2e-5, atol=1e-6), name
        for a, b in [(c.base_weight, o.base_weight), (c.left_weight, o.left_weight), (c.right_weight, o.right_weight)]:
            assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32), name


def test_categorical_evaluation_with_a_thousand_categories(xb):
    """One feature with 1000 categories: the categorical kernels' sort scratch (8 warps x 1000 bins x 6 B) plus their
    static shared memory passes 48 KiB, so the launch needs the per-kernel opt-in (AllowDynamicShared); the sorted
    partition spans 32 chunks of 32 bins."""
    import torch
    rng = np.random.default_rng(41)
    n, k = 60000, 1000
    col = rng.integers(0, k, size=n).astype(np.float32)
    col[rng.random(n) < 0.05] = np.nan
    xnum, _ = make_regression(n, 1, seed=29)
    ncp, ncv = make_cuts(xnum, 32)
    cp = np.array([0, k, k + int(ncp[1])], dtype=np.uint32)
    cv = np.array([float(v) for v in range(k)] + [float(v) for v in ncv], dtype=np.float32)
    is_cat = [1, 0]
    gb = np.full((n, 2), -1, dtype=np.int32)
    ok = ~np.isnan(col)
    gb[ok, 0] = col[ok].astype(np.int32)
    gb[:, 1] = orc.bin_dense(xnum, ncp, ncv)[:, 0] + k
    effect = rng.normal(size=k)
    y = (effect[np.nan_to_num(col, nan=0).astype(int)] + 0.2 * rng.normal(size=n)).astype(np.float32)
    gp = squared_error_gpair(np.zeros(n, np.float32), y)
    scale, _ = orc.quantiser(gp)
    gq = orc.quantise(gp, scale)
    subsets = [np.arange(n), rng.choice(n, 5000, replace=False)]
    hists = np.stack([orc.build_hist(gb, gq, s.astype(np.uint32), cv.size) for s in subsets])
    parents = [tuple(int(v) for v in gq[s].sum(axis=0)) for s in subsets]
    q = xb.lib.Quantiser(*[float(v) for v in scale])
    page = xb.EllpackPage(xb.HistogramCuts(cp, cv), 1, 0, False)
    for threshold in (64, 1000):
        params = dict(max_depth=6)
        cands, flags, bits = xb.components.evaluate_splits_cat(
            page, torch.from_numpy(hists).cuda(), parents, xb.TrainParam.make(**params), q, is_cat,
            max_cat_to_onehot=4, max_cat_threshold=threshold, n_cat_words=(k + 31) // 32)
        for j, c in enumerate(cands):
            o, o_is_cat, _, o_cats = orc.eval_node_cat(hists[j], parents[j], cp, cv, orc.Param.make(**params), scale,
                                                       is_cat, 4, threshold)
            assert o_is_cat and bool(flags[j]), (threshold, j)
            assert (c.findex, c.dir) == (o.findex, o.dir), (threshold, j)
            assert _cats_of(bits[j]) == o_cats, (threshold, j)
            assert (c.left_sum.grad, c.left_sum.hess, c.right_sum.grad, c.right_sum.hess) == \
                (o.left_sum.g, o.left_sum.h, o.right_sum.g, o.right_sum.h), (threshold, j)


# ---------------------------------------------------------------------------------------------------------------------
# Categorical features through the whole updater (b200hist_updater_set_categorical: evaluation, ApplyKernel<true>,
# PartitionCountKernel<., true>, FinaliseWindowKernel<., true>)
# ---------------------------------------------------------------------------------------------------------------------
def _cat_page(xb, x, cp, cv, is_cat):
    """The page's binning kernel searches numerically (first cut greater than the value); a category code c binned
    that way lands one bin late, c - 0.5 lands in its own bin (codes are 0..K-1 here).  Proper SearchCatBin binning at
    page construction is not built yet: the drop-in imports pages the reference has binned."""
    xs = x.copy()
    for j, c in enumerate(is_cat):
        if c:
            xs[:, j] -= 0.5
    return _page(xb, xs, cp, cv)


def _compare_cat_tree(xb, up, tree, o, cp, cv):
    """_compare_tree, but a categorical split's condition is NaN on both sides and its category set is compared."""
    nodes = o["nodes"]
    assert tree.num_nodes == len(nodes)
    is_cat, sets = up.last_tree_categories(tree.num_nodes)
    for i, (a, b) in enumerate(zip(tree.nodes, nodes)):
        assert (a["parent"], a["left"], a["right"], a["is_leaf"], a["depth"]) == \
            (b.parent, b.left, b.right, bool(b.is_leaf), b.depth), i
        assert int(is_cat[i]) == int(o["is_cat"][i]), i
        if not b.is_leaf:
            assert (a["findex"], a["default_left"]) == (b.findex, bool(b.default_left)), i
            if o["is_cat"][i]:
                assert np.isnan(a["fvalue"]) and sets[i] == o["cat_sets"][i], (i, sets[i], o["cat_sets"][i])
            else:
                assert np.float32(a["fvalue"]).view(np.uint32) == np.float32(b.fvalue).view(np.uint32), i
            # the device divides with div.approx.ftz.f32 like the reference, the oracle in IEEE arithmetic: the error
            # of loss_chg = split gain - parent gain scales with the gains, not with their difference
            # (same bound as tests/test_gpu_parity._compare_tree)
            c = o["cands"][i]
            pg = float(c.left_sum.g + c.right_sum.g) * o["scale"][2]
            ph = float(c.left_sum.h + c.right_sum.h) * o["scale"][3]
            gain_scale = abs(b.loss_chg) + pg * pg / max(ph, 1e-30)
            assert abs(a["loss_chg"] - b.loss_chg) <= 2e-5 * gain_scale + 1e-6, i
        else:
            assert np.float32(a["leaf_value"]).view(np.uint32) == np.float32(b.leaf_value).view(np.uint32), i
        assert np.float32(a["sum_hess"]).view(np.uint32) == np.float32(b.sum_hess).view(np.uint32), i
        assert np.float32(a["base_weight"]).view(np.uint32) == np.float32(b.base_weight).view(np.uint32), i
    for i in range(len(nodes)):
        if o["hist"] is not None and o["hist"][i].any():
            assert np.array_equal(up.node_histogram(i), o["hist"][i]), "hist of node %d" % i


@pytest.mark.parametrize("case", json.load(open(os.path.join(GOLDEN, "ref_cpu_categorical_trees.json")))["cases"],
                         ids=lambda c: c["name"])
def test_categorical_trees_match_oracle_and_cpu_reference(xb, case):
    """Data and cuts of the CPU reference fixture: our tree equals the oracle's node for node (category sets,
    histograms, positions), and with it the discrete content of the reference's CPU tree."""
    import torch
    from tests.test_reference_golden import categorical_global_bins
    from tests.helpers import make_categorical_data as make_data
    x, y, types = make_data(case)
    is_cat = [1 if t == "c" else 0 for t in types]
    cp = np.array(case["cut_ptrs"], dtype=np.uint32)
    cv = np.array(case["cut_values_bits"], dtype=np.uint32).view(np.float32)
    gb = categorical_global_bins(x, cp, cv, is_cat)
    page = _cat_page(xb, x, cp, cv, is_cat)
    assert np.array_equal(_global_bins(page, cp), gb), "test binning recipe"
    params = dict(case["params"])
    up = xb.GPUHistMaker(page)
    up.configure(params)
    up.set_categorical(is_cat, case["max_cat_to_onehot"], case["max_cat_threshold"])
    pred = np.zeros(case["n"], dtype=np.float32)
    pos = torch.empty(case["n"], dtype=torch.int32, device="cuda")
    for it in range(2):
        gp = squared_error_gpair(pred, y)
        tree = up.update(torch.from_numpy(gp).cuda(), pos)
        o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), want_hist=True, feature_is_cat=is_cat,
                            max_cat_to_onehot=case["max_cat_to_onehot"], max_cat_threshold=case["max_cat_threshold"])
        _compare_cat_tree(xb, up, tree, o, cp, cv)
        assert np.array_equal(pos.cpu().numpy(), o["position"])
        if it == 0:  # the fixture's tree
            ref = case["tree"]
            for i, nd in enumerate(tree.nodes):
                if not nd["is_leaf"]:
                    assert nd["findex"] == ref["findex"][i], i
        pred = orc.update_prediction_cache(o["nodes"], o["position"], pred)
    up.set_categorical(None)
    gp = squared_error_gpair(np.zeros(case["n"], np.float32), y)
    t_plain = up.update(torch.from_numpy(gp).cuda())
    o_plain = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params))
    _compare_tree(xb, up, t_plain, o_plain, cp, cv, hists=False)


@pytest.mark.parametrize("extra", [dict(max_depth=6), dict(max_depth=0, max_leaves=20, grow_policy="lossguide"),
                                   dict(max_depth=5, subsample=0.7), dict(max_depth=5, mono=True)],
                         ids=["depthwise", "lossguide", "subsample", "monotone"])
def test_categorical_trees_with_other_parameters(xb, extra):
    """70 categories (more than two 32-bin chunks, uint8 page), sampling and monotone constraints next to categorical
    features, loss-guided growth."""
    import torch
    rng = np.random.default_rng(91)
    n = 15000
    ks = [70, 5, 33]
    cols = [rng.integers(0, k, size=n).astype(np.float32) for k in ks]
    cols[0][rng.random(n) < 0.1] = np.nan
    xn, _ = make_regression(n, 3, seed=92, missing=0.05)
    x = np.column_stack(cols + [xn[:, j] for j in range(3)]).astype(np.float32)
    is_cat = [1, 1, 1, 0, 0, 0]
    ncp, ncv = make_cuts(xn, 32)
    cut_values, cut_ptrs = [], [0]
    for k in ks:
        cut_values += [float(v) for v in range(k)]
        cut_ptrs.append(len(cut_values))
    for j in range(3):
        cut_values += [float(v) for v in ncv[ncp[j]:ncp[j + 1]]]
        cut_ptrs.append(len(cut_values))
    cp, cv = np.array(cut_ptrs, dtype=np.uint32), np.array(cut_values, dtype=np.float32)
    from tests.test_reference_golden import categorical_global_bins
    gb = categorical_global_bins(x, cp, cv, is_cat)
    eff = [rng.normal(size=k) for k in ks]
    y = (sum(e[np.nan_to_num(c, nan=0).astype(int)] for e, c in zip(eff, cols)) + np.nan_to_num(xn[:, 0]) +
         0.2 * rng.normal(size=n)).astype(np.float32)
    params = {k: v for k, v in extra.items() if k not in ("subsample", "mono")}
    params["eta"] = 0.5
    page = _cat_page(xb, x, cp, cv, is_cat)
    up = xb.GPUHistMaker(page)
    up.configure(params)
    up.set_categorical(is_cat, 4, 64)
    mono = [0, 0, 0, 1, -1, 0] if extra.get("mono") else None
    if mono:
        up.set_monotone_constraints(mono)
    sub = extra.get("subsample", 1.0)
    pred = np.zeros(n, dtype=np.float32)
    pos = torch.empty(n, dtype=torch.int32, device="cuda")
    for it in range(3):
        gp = squared_error_gpair(pred, y)
        if sub < 1:
            up.set_tree_sampling(None, sub, 1000 + it)
        tree = up.update(torch.from_numpy(gp).cuda(), pos)
        o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), want_hist=True, feature_is_cat=is_cat,
                            subsample=sub, sample_seed=1000 + it, monotone=mono)
        _compare_cat_tree(xb, up, tree, o, cp, cv)
        assert np.array_equal(pos.cpu().numpy(), o["position"])
        assert o["is_cat"].sum() >= 1
        pred = orc.update_prediction_cache(o["nodes"], o["position"], pred)


# ---------------------------------------------------------------------------------------------------------------------
# Drop-in (the reference's own Booster, our factory): interaction constraints and categorical features through the
# shim.
# ---------------------------------------------------------------------------------------------------------------------
def _shim_or_skip():
    import ctypes
    from oracle import ref_capi as rc
    from tests import test_shim_dropin as sd
    if not (rc.available("cuda") and os.path.exists(sd.SHIM)):
        pytest.skip("needs oracle/_ref_cuda/libxgboost.so and the shim")
    rc.load("cuda")
    f = sd.test_reference_booster_trains_identically_through_the_shim
    if getattr(f, "_shim", None) is None:
        f._shim = ctypes.CDLL(sd.SHIM, mode=ctypes.RTLD_GLOBAL)
    return rc


def _model_trees(bst):
    keep = ("left_children", "right_children", "split_indices", "split_conditions", "default_left", "base_weights",
            "loss_changes", "sum_hessian", "split_type", "categories", "categories_nodes", "categories_segments",
            "categories_sizes")
    return [{k: t[k] for k in keep} for t in bst.save_json()["learner"]["gradient_booster"]["model"]["trees"]]


@pytest.mark.parametrize("extra", [
    {"interaction_constraints": "[[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [5, 6]]"},
    {"interaction_constraints": "[[0, 1, 2], [3, 4, 5]]", "subsample": 0.8},
    {"interaction_constraints": "[[0, 1], [1, 2], [2, 3], [3, 4]]", "colsample_bylevel": 0.7},
    {"interaction_constraints": "[[0, 5], [5, 9], [9, 2]]", "grow_policy": "lossguide", "max_leaves": 24, "max_depth": 0},
], ids=["chain", "disjoint_subsample", "chain_bylevel", "lossguide"])
def test_interaction_constraints_through_the_shim_match_reference(xb, extra):
    import torch
    rc = _shim_or_skip()
    x, y = make_regression(30000, 12, seed=81)
    y = (y + x[:, 0] * x[:, 1] + x[:, 2] * x[:, 3] - x[:, 4] * x[:, 5]).astype(np.float32)
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": 64, "max_depth": 6, "eta": 0.3,
              "objective": "reg:squarederror", "base_score": 0.5}
    params.update(extra)

    def train(updater):
        dm = rc.DMatrix.from_cuda(torch.from_numpy(x).cuda(), kind="cuda")
        dm.set_label(y)
        bst = rc.Booster(dict(params, updater=updater), dm)
        for i in range(4):
            bst.update(i)
        return _model_trees(bst), bst.predict_margin(dm)
    ref_trees, ref_margin = train("grow_gpu_hist_reference")
    our_trees, our_margin = train("grow_b200_hist")
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))


@pytest.mark.parametrize("extra", [{}, {"max_cat_to_onehot": 1}, {"max_cat_to_onehot": 100},
                                   {"max_cat_threshold": 8}, {"subsample": 0.7, "colsample_bytree": 0.7},
                                   {"grow_policy": "lossguide", "max_leaves": 24, "max_depth": 0}],
                         ids=lambda e: "_".join("%s%s" % kv for kv in e.items()) or "defaults")
def test_categorical_features_through_the_shim_match_reference(xb, extra):
    """The reference sketches, bins and stores the categorical columns; our updater grows the trees.  Models must
    be identical including split_type / categories arrays, and so must the margins (the reference's predictor walks
    OUR category sets)."""
    from tests.helpers import make_categorical_data as make_data
    rc = _shim_or_skip()
    x, y, types = make_data(dict(seed=305, n=30000))
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": 64, "max_depth": 6, "eta": 0.3,
              "objective": "reg:squarederror", "base_score": 0.5}
    params.update(extra)

    def train(updater):
        dm = rc.DMatrix.from_dense(x, label=y, kind="cuda")
        dm.set_feature_types(types)
        bst = rc.Booster(dict(params, updater=updater), dm)
        for i in range(4):
            bst.update(i)
        return _model_trees(bst), bst.predict_margin(dm)
    ref_trees, ref_margin = train("grow_gpu_hist_reference")
    our_trees, our_margin = train("grow_b200_hist")
    assert sum(sum(t["split_type"]) for t in ref_trees) >= 4, "the fixture must produce categorical splits"
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))


# ---------------------------------------------------------------------------------------------------------------------
# colsample_bynode (b200hist_updater_set_node_feature_sets, NodeSetLevelKernel)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("extra", [dict(max_depth=6), dict(max_depth=0, max_leaves=20, grow_policy="lossguide"),
                                   dict(max_depth=5, mono=True), dict(max_depth=7, max_leaves=17)],
                         ids=["depthwise", "lossguide", "monotone", "depthwise_max_leaves"])
def test_node_feature_sets_match_oracle(xb, extra):
    import torch
    n, f = 15000, 24
    x, y = make_regression(n, f, seed=71, missing=0.05)
    cp, cv = make_cuts(x, 32)
    page = _page(xb, x, cp, cv)
    gb = orc.bin_dense(x, cp, cv)
    params = {k: v for k, v in extra.items() if k != "mono"}
    params["eta"] = 0.5
    up = xb.GPUHistMaker(page)
    up.configure(params)
    mono = ([1, -1, 0] * 8) if extra.get("mono") else None
    if mono:
        up.set_monotone_constraints(mono)
    rng = np.random.default_rng(72)
    pred = np.zeros(n, dtype=np.float32)
    for it in range(3):
        sets = np.stack([np.sort(rng.permutation(f)[:9]) for _ in range(up.max_nodes)]).astype(np.int32)
        up.set_node_feature_sets(sets)
        gp = squared_error_gpair(pred, y)
        tree = up.update(torch.from_numpy(gp).cuda())
        o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), want_hist=True, node_feature_sets=sets,
                            monotone=mono)
        _compare_tree(xb, up, tree, o, cp, cv, hists=True)
        assert up.last_eval_count() == o["eval_count"]
        pred = orc.update_prediction_cache(o["nodes"], o["position"], pred)
    up.set_node_feature_sets(None)
    gp = squared_error_gpair(np.zeros(n, np.float32), y)
    t_plain = up.update(torch.from_numpy(gp).cuda())
    o_plain = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), monotone=mono)
    _compare_tree(xb, up, t_plain, o_plain, cp, cv, hists=False)


@pytest.mark.parametrize("extra", [{"colsample_bynode": 0.5}, {"colsample_bynode": 0.6, "colsample_bytree": 0.7, "subsample": 0.8},
                                   {"colsample_bynode": 0.4, "grow_policy": "lossguide", "max_leaves": 24, "max_depth": 0}],
                         ids=lambda e: "_".join("%s%s" % kv for kv in e.items()))
def test_colsample_bynode_through_the_shim_matches_reference(xb, extra):
    """The shim draws the tree's whole stream of node sets with the reference's own ColumnSampler from a copy of the
    Booster's engine and replays the engine by the number of nodes evaluated: models and the engine state after
    every tree (hence all later trees) must equal the reference's."""
    import torch
    rc = _shim_or_skip()
    x, y = make_regression(30000, 40, seed=83)
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": 64, "max_depth": 5, "eta": 0.3,
              "objective": "reg:squarederror", "base_score": 0.5}
    params.update(extra)

    def train(updater):
        dm = rc.DMatrix.from_cuda(torch.from_numpy(x).cuda(), kind="cuda")
        dm.set_label(y)
        bst = rc.Booster(dict(params, updater=updater), dm)
        for i in range(4):
            bst.update(i)
        return _model_trees(bst), bst.predict_margin(dm)
    ref_trees, ref_margin = train("grow_gpu_hist_reference")
    our_trees, our_margin = train("grow_b200_hist")
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))


# ---------------------------------------------------------------------------------------------------------------------
# Categorical columns at DMatrix construction: cut values (b200hist_sketch_set_feature_types) and SearchCatBin binning
# (b200hist_ellpack_set_feature_types, RebinCategoricalKernel)
# ---------------------------------------------------------------------------------------------------------------------
def test_categorical_cuts_and_binning(xb):
    import torch
    from tests.test_reference_golden import categorical_global_bins
    from tests.helpers import CATEGORICAL_N_CATS as N_CATS, make_categorical_data as make_data
    case = json.load(open(os.path.join(GOLDEN, "ref_cpu_categorical_trees.json")))["cases"][0]
    x, _, types = make_data(case)
    is_cat = [1 if t == "c" else 0 for t in types]
    xd = torch.from_numpy(x).cuda()
    sk = xb.QuantileSketch(x.shape[1], case["max_bin"], torch.device("cuda", 0))
    sk.set_feature_types(is_cat)
    for b in range(0, x.shape[0], 3000):          # several windows
        sk.push(xd[b:b + 3000])
    cuts = sk.make_cuts()
    sk.close()
    ref_cp = np.array(case["cut_ptrs"], dtype=np.uint32)
    ref_cv = np.array(case["cut_values_bits"], dtype=np.uint32).view(np.float32)
    for j, k in enumerate(N_CATS):                 # the reference's categorical cuts: every code 0 .. max
        seg = cuts.cut_values[cuts.cut_ptrs[j]:cuts.cut_ptrs[j + 1]]
        assert np.array_equal(seg, np.arange(k, dtype=np.float32)), j
        assert np.array_equal(seg, ref_cv[ref_cp[j]:ref_cp[j + 1]]), j
    # numerical columns are untouched by the flags
    plain = xb.QuantileSketch(x.shape[1], case["max_bin"], torch.device("cuda", 0))
    for b in range(0, x.shape[0], 3000):
        plain.push(xd[b:b + 3000])
    pc = plain.make_cuts()
    plain.close()
    for j in range(len(N_CATS), x.shape[1]):
        assert np.array_equal(cuts.cut_values[cuts.cut_ptrs[j]:cuts.cut_ptrs[j + 1]],
                              pc.cut_values[pc.cut_ptrs[j]:pc.cut_ptrs[j + 1]]), j
    # binning: a category lands in the bin of its code, a value between codes truncates, too large clamps
    page = xb.EllpackPage.from_dense(xd, cuts, feature_is_cat=is_cat)
    gb = categorical_global_bins(x, cuts.cut_ptrs, cuts.cut_values, is_cat)
    assert np.array_equal(_global_bins(page, cuts.cut_ptrs), gb)
    odd = x[:64].copy()
    odd[:, 0] = np.array([0.0, 0.9, 1.2, 2.0, 2.7, 9.0, np.nan, 1.0] * 8, dtype=np.float32)
    page2 = xb.EllpackPage.from_dense(torch.from_numpy(odd).cuda(), cuts, has_missing=True, feature_is_cat=is_cat)
    want = np.array([0, 0, 1, 2, 2, 2, -1, 1] * 8)
    assert np.array_equal(page2.symbols()[:, 0], want)
    # an invalid category is an error, as in the reference
    bad = xb.QuantileSketch(x.shape[1], case["max_bin"], torch.device("cuda", 0))
    bad.set_feature_types(is_cat)
    neg = x[:100].copy()
    neg[3, 1] = -1.0
    bad.push(torch.from_numpy(neg).cuda())
    with pytest.raises(xb.B200HistError):
        bad.make_cuts()
    bad.close()


# ---------------------------------------------------------------------------------------------------------------------
# Prediction on quantised pages (b200hist_forest_*, csrc/predict.cu)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("missing,max_bin", [(0.0, 64), (0.1, 256), (0.05, 600)], ids=["dense", "missing_u16", "wide_bins"])
def test_forest_prediction_matches_oracle(xb, missing, max_bin):
    """Train a few trees, then predict the training page and a VALIDATION page quantised with the same cuts: margins
    bit for bit (tree-order float sum from 0, then added to the output) and leaf ids, against the oracle's walk."""
    import torch
    n, f = 20000, 20
    x, y = make_regression(n, f, seed=95, missing=missing)
    xv, _ = make_regression(7000, f, seed=96, missing=missing)
    cp, cv = make_cuts(x, max_bin)
    page = _page(xb, x, cp, cv, has_missing=missing > 0)
    vpage = _page(xb, xv, cp, cv, has_missing=missing > 0)
    gb, gbv = orc.bin_dense(x, cp, cv), orc.bin_dense(xv, cp, cv)
    params = dict(max_depth=6, eta=0.3)
    up = xb.GPUHistMaker(page)
    up.configure(params)
    forest = xb.Forest(page)
    pred = np.zeros(n, dtype=np.float32)
    trees = []
    for it in range(6):
        gp = squared_error_gpair(pred, y)
        tree = up.update(torch.from_numpy(gp).cuda())
        forest.add_tree(tree)
        o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params))
        trees.append(o["nodes"])
        pred = orc.update_prediction_cache(o["nodes"], o["position"], pred)
    assert forest.num_trees == 6
    for pg, bins, rows in ((page, gb, n), (vpage, gbv, 7000)):
        base = np.full(rows, 0.5, dtype=np.float32)
        out = forest.predict_margin(pg, torch.from_numpy(base.copy()).cuda())
        want = orc.predict_margin(bins, cp, cv, trees, base=base)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
        leaves = forest.predict_leaf(pg).cpu().numpy()
        for t, nodes in enumerate(trees):
            assert np.array_equal(leaves[:, t], orc.predict_leaf(bins, cp, cv, nodes))
        assert np.array_equal(forest.predict_leaf(pg, as_float=True).cpu().numpy(), leaves.astype(np.float32))
        # three output groups: leaves straight onto the output, tree after tree
        groups = [1, 0, 2, 2, 0, 1]
        forest.set_tree_groups(groups)
        gbase = np.random.default_rng(rows).normal(size=(rows, 3)).astype(np.float32)
        gout = forest.predict_margin_groups(pg, torch.from_numpy(gbase.copy()).cuda(), 3)
        gwant = orc.predict_margin_groups(bins, cp, cv, trees, groups, 3, base=gbase)
        assert np.array_equal(gout.cpu().numpy().view(np.uint32), gwant.view(np.uint32))
        # the same three results straight from the page as the reference stores it (bit-packed gidx_buffer)
        from tests.helpers import pack_reference_page
        buf, stride, n_symbols, is_dense = pack_reference_page(bins, cp, dense_layout=True)
        dbuf = torch.from_numpy(buf).cuda()
        r0 = forest.predict_reference_page(dbuf, rows, stride, n_symbols, True, is_dense,
                                           torch.from_numpy(base.copy()).cuda(), mode=0)
        assert np.array_equal(r0.cpu().numpy().view(np.uint32), want.view(np.uint32))
        r1 = forest.predict_reference_page(dbuf, rows, stride, n_symbols, True, is_dense,
                                           torch.from_numpy(gbase.copy()).cuda(), mode=1, n_groups=3)
        assert np.array_equal(r1.cpu().numpy().view(np.uint32), gwant.view(np.uint32))
        r2 = forest.predict_reference_page(dbuf, rows, stride, n_symbols, True, is_dense,
                                           torch.empty((rows, 6), dtype=torch.float32, device="cuda"), mode=2)
        assert np.array_equal(r2.cpu().numpy(), leaves.astype(np.float32))
        if missing > 0:   # and from a sparse-layout page (ascending global bins, searched per row)
            sbuf, sstride, ssym, _ = pack_reference_page(bins, cp, dense_layout=False)
            r3 = forest.predict_reference_page(torch.from_numpy(sbuf).cuda(), rows, sstride, ssym, False, False,
                                               torch.from_numpy(base.copy()).cuda(), mode=0)
            assert np.array_equal(r3.cpu().numpy().view(np.uint32), want.view(np.uint32))
        forest.set_tree_groups([0] * 6)
        # a tree range and a row range
        part = forest.predict_margin(pg, torch.zeros(1000, device="cuda"), tree_begin=2, tree_end=5, row_begin=123)
        want_part = orc.predict_margin(bins[123:1123], cp, cv, trees[2:5])
        assert np.array_equal(part.cpu().numpy().view(np.uint32), want_part.view(np.uint32))
    # a page with other cuts is refused
    cp2, cv2 = make_cuts(xv, max_bin)
    other = _page(xb, xv, cp2, cv2, has_missing=missing > 0)
    with pytest.raises(xb.B200HistError):
        forest.predict_margin(other, torch.zeros(7000, device="cuda"))


def _as_tree(nodes):
    """An oracle tree (list of orc.Node) as the RegTree Forest.add_tree takes: the two structs have the same fields."""
    from xgboost_b200 import lib as _l
    from xgboost_b200.updater import RegTree
    return RegTree([_l.TreeNode(*[getattr(nd, name) for name, _ in _l.TreeNode._fields_]) for nd in nodes])


def test_forest_resolved_against_foreign_cuts_matches_oracle(xb):
    """A forest grown elsewhere (other cuts), resolved against the cuts of the page it predicts: the conditions are
    not cuts of that page, and the reference's loader rule (the bin's lower cut, gpu_data_accessor.cuh:95-107)
    decides -- the oracle's walk is pinned on the reference for exactly this case
    (tests/test_reference_golden.py::test_oracle_predictor_on_foreign_cuts_...)."""
    import torch
    x, y = make_regression(15000, 12, seed=97, missing=0.05)
    cp, cv = make_cuts(x, 64)
    gb = orc.bin_dense(x, cp, cv)
    pred = np.zeros(len(y), dtype=np.float32)
    trees = []
    for _ in range(5):
        o = orc.update_tree(gb, cp, cv, squared_error_gpair(pred, y), orc.Param.make(max_depth=5, eta=0.3))
        trees.append(o["nodes"])
        pred = orc.update_prediction_cache(o["nodes"], o["position"], pred)
    xv, _ = make_regression(9000, 12, seed=98, missing=0.05)
    cp2, cv2 = make_cuts(xv, 24)
    vpage = _page(xb, xv, cp2, cv2, has_missing=True)
    gbv = orc.bin_dense(xv, cp2, cv2)
    forest = xb.Forest(vpage)
    for nodes in trees:
        forest.add_tree(_as_tree(nodes))
    out = forest.predict_margin(vpage, torch.zeros(9000, device="cuda"))
    want = orc.predict_margin(gbv, cp2, cv2, trees)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("case", ["dense", "missing", "wide_bins", "multiclass"])
def test_predictor_binding_serves_quantised_matrices_like_the_reference(xb, case, tmp_path):
    """The reference's Booster.predict on QuantileDMatrices with the "gpu_predictor" factory served by the shim
    (xgboost_b200/csrc/shim/b200_predictor.cu) against the same call served by the reference itself: margins bit for
    bit (one output or one per class), on the training matrix and on a matrix with its own cuts, leaf ids too, and the
    shim's run really launched our kernels.
    Two child processes: the shim decides at load time (B200HIST_SHIM_PREDICTOR)."""
    import subprocess
    import sys
    from oracle import ref_capi as rc
    from tests import test_shim_dropin as sd
    from tests.conftest import ROOT
    if not (rc.available("cuda") and os.path.exists(sd.SHIM)):
        pytest.skip("needs oracle/_ref_cuda/libxgboost.so and the shim")
    got = {}
    for gate in ("0", "1"):
        out = os.path.join(str(tmp_path), "margins_%s.npz" % gate)
        env = dict(os.environ, B200HIST_SHIM_PREDICTOR=gate)
        subprocess.run([sys.executable, "-m", "tests.shim_predictor_check", out, case], cwd=ROOT, env=env, check=True,
                       timeout=600)
        got[gate] = np.load(out)
    assert int(got["0"]["launches"]) == 0
    # margins of the validation matrix + leaf ids: two predict launches at least (the training matrix's margins come
    # from the Booster's prediction cache, no predictor call)
    assert int(got["1"]["launches"]) >= 2
    for key in ("valid", "train", "leaf"):
        assert np.array_equal(got["1"][key].view(np.uint32), got["0"][key].view(np.uint32)), key


def test_forest_prediction_with_categorical_splits(xb):
    import torch
    from tests.test_reference_golden import categorical_global_bins
    from tests.helpers import make_categorical_data as make_data
    case = json.load(open(os.path.join(GOLDEN, "ref_cpu_categorical_trees.json")))["cases"][0]
    x, y, types = make_data(case)
    is_cat = [1 if t == "c" else 0 for t in types]
    cp = np.array(case["cut_ptrs"], dtype=np.uint32)
    cv = np.array(case["cut_values_bits"], dtype=np.uint32).view(np.float32)
    gb = categorical_global_bins(x, cp, cv, is_cat)
    page = _cat_page(xb, x, cp, cv, is_cat)
    up = xb.GPUHistMaker(page)
    up.configure(dict(case["params"]))
    up.set_categorical(is_cat, case["max_cat_to_onehot"], case["max_cat_threshold"])
    forest = xb.Forest(page)
    pred = np.zeros(case["n"], dtype=np.float32)
    trees = []
    for it in range(3):
        gp = squared_error_gpair(pred, y)
        tree = up.update(torch.from_numpy(gp).cuda())
        flags, sets = up.last_tree_categories(tree.num_nodes)
        forest.add_tree(tree, flags, sets)
        o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**case["params"]), feature_is_cat=is_cat,
                            max_cat_to_onehot=case["max_cat_to_onehot"], max_cat_threshold=case["max_cat_threshold"])
        trees.append(o)
        pred = orc.update_prediction_cache(o["nodes"], o["position"], pred)
    out = forest.predict_margin(page, torch.zeros(case["n"], device="cuda"))
    want = orc.predict_margin(gb, cp, cv, trees)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert np.array_equal(want.view(np.uint32), pred.view(np.uint32)) or np.allclose(want, pred, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------------
# Gradient-based row sampling (b200hist_updater_set_sampling_method, csrc/sampling.cu).  Its oracle is itself waiting
# for a GPU-reference fixture (tests/golden/ref_gpu_next_trees.json): until then this compares two restatements.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("subsample,seed,obj", [(0.3, 12345, "sq"), (0.7, 4294967295, "sq"), (0.1, 7, "logistic"),
                                                (0.9, 0, "sq")])
def test_gradient_based_sampling_matches_oracle(xb, subsample, seed, obj):
    import torch
    from tests.helpers import logistic_gpair
    n, f = 30000, 12
    x, y = make_regression(n, f, seed=97)
    if obj == "logistic":
        y = (y > 0).astype(np.float32)
    cp, cv = make_cuts(x, 64)
    page = _page(xb, x, cp, cv)
    gb = orc.bin_dense(x, cp, cv)
    params = dict(max_depth=5, eta=0.3)
    up = xb.GPUHistMaker(page)
    up.configure(params)
    up.set_sampling_method("gradient_based")
    pred = np.zeros(n, dtype=np.float32)
    pos = torch.empty(n, dtype=torch.int32, device="cuda")
    for it in range(3):
        gp = squared_error_gpair(pred, y) if obj == "sq" else logistic_gpair(pred, y)
        up.set_tree_sampling(None, subsample, seed + it)
        tree = up.update(torch.from_numpy(gp).cuda(), pos)
        o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), want_hist=True, subsample=subsample,
                            sample_seed=(seed + it) & 0xFFFFFFFF, sampling_method="gradient_based")
        _compare_tree(xb, up, tree, o, cp, cv, hists=True)
        assert np.array_equal(pos.cpu().numpy(), o["position"])       # ~leaf for the rows that lost their trial
        dropped = float((o["position"] < 0).mean())
        assert abs((1.0 - dropped) - subsample) < 0.05
        pred = orc.update_prediction_cache(o["nodes"], o["position"], pred)
    up.set_sampling_method("uniform")
    up.set_tree_sampling(None, subsample, seed)
    gp = squared_error_gpair(np.zeros(n, np.float32), y)
    t_uni = up.update(torch.from_numpy(gp).cuda())
    o_uni = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), subsample=subsample, sample_seed=seed & 0xFFFFFFFF)
    _compare_tree(xb, up, t_uni, o_uni, cp, cv, hists=False)


@pytest.mark.parametrize("extra", [{"subsample": 0.3}, {"subsample": 0.6, "colsample_bytree": 0.7, "seed": 5}],
                         ids=lambda e: "_".join("%s%s" % kv for kv in e.items()))
def test_gradient_based_sampling_through_the_shim_matches_reference(xb, extra):
    import torch
    rc = _shim_or_skip()
    x, y = make_regression(40000, 20, seed=98)
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": 64, "max_depth": 5, "eta": 0.3,
              "objective": "reg:squarederror", "base_score": 0.5, "sampling_method": "gradient_based"}
    params.update(extra)

    def train(updater):
        dm = rc.DMatrix.from_cuda(torch.from_numpy(x).cuda(), kind="cuda")
        dm.set_label(y)
        bst = rc.Booster(dict(params, updater=updater), dm)
        for i in range(4):
            bst.update(i)
        return _model_trees(bst), bst.predict_margin(dm)
    ref_trees, ref_margin = train("grow_gpu_hist_reference")
    our_trees, our_margin = train("grow_b200_hist")
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))


# ---------------------------------------------------------------------------------------------------------------------
# More of the reference's callers through the (already validated) default path: several trees per Update call.
# Gated only because they have not run yet; nothing new on the device side.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("extra", [{"objective": "multi:softprob", "num_class": 3},
                                   {"num_parallel_tree": 3, "subsample": 0.7, "colsample_bynode": 1.0},
                                   {"objective": "multi:softmax", "num_class": 4, "colsample_bytree": 0.6, "subsample": 0.8}],
                         ids=["multiclass", "random_forest", "multiclass_sampled"])
def test_several_trees_per_iteration_through_the_shim(xb, extra):
    """Multi-class boosting calls the updater once per class with that class's gradient column; num_parallel_tree
    hands several trees to one Update call, each with its own sampling draws."""
    import torch
    rc = _shim_or_skip()
    x, y = make_regression(30000, 16, seed=99)
    if "num_class" in extra:
        y = np.digitize(y, np.quantile(y, np.linspace(0, 1, extra["num_class"] + 1)[1:-1])).astype(np.float32)
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": 64, "max_depth": 5, "eta": 0.3,
              "objective": "reg:squarederror", "base_score": 0.5}
    params.update(extra)

    def train(updater):
        dm = rc.DMatrix.from_cuda(torch.from_numpy(x).cuda(), kind="cuda")
        dm.set_label(y)
        bst = rc.Booster(dict(params, updater=updater), dm)
        for i in range(3):
            bst.update(i)
        return _model_trees(bst)
    ref_trees = train("grow_gpu_hist_reference")
    our_trees = train("grow_b200_hist")
    assert len(ref_trees) == 3 * max(extra.get("num_class", 1), extra.get("num_parallel_tree", 1))
    assert json.dumps(our_trees) == json.dumps(ref_trees)


# ---------------------------------------------------------------------------------------------------------------------
# Column-sampling draws on the device (b200hist_col_sample_stream, csrc/colsample.cu)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,colsample", [(1, 0.5), (7, 0.5), (16, 0.3), (17, 0.9), (256, 0.5), (257, 0.25), (1000, 0.01),
                                         (4096, 0.7), (8192, 0.5), (8192, 0.9), (40, 1.0)])
def test_col_sample_stream_matches_the_thrust_shuffle_restatement(xb, n, colsample):
    """Every draw of the stream equals ColSample restated on the CPU (oracle.col_sample, itself pinned on the toolkit's
    thrust::shuffle): sizes around the cipher's power-of-two boundaries, the engine's seed folding, a sparse id set."""
    rng = np.random.default_rng(n)
    features = np.sort(rng.choice(3 * n + 5, size=n, replace=False)).astype(np.int32)
    seeds = [0, 1, 2147483647, 4294967295] + [int(v) for v in rng.integers(0, 2 ** 32, size=60)]
    got = xb.components.col_sample_stream(features, colsample, seeds)
    assert got.shape[0] == len(seeds)
    for d, seed in enumerate(seeds):
        assert np.array_equal(got[d], orc.col_sample(features, colsample, seed)), (n, colsample, seed)


# ---------------------------------------------------------------------------------------------------------------------
# The seed-driven column sampler on the device (b200hist_updater_set_column_sampler: LevelDrawKernel / NodeDrawKernel)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bylevel,bynode,extra", [
    (1.0, 0.5, dict(max_depth=6)), (0.6, 0.5, dict(max_depth=6)), (0.6, 1.0, dict(max_depth=5)),
    (0.7, 0.6, dict(max_depth=0, max_leaves=20, grow_policy="lossguide")), (0.5, 0.5, dict(max_depth=7, max_leaves=17))],
    ids=["bynode", "bylevel_bynode", "bylevel", "lossguide", "max_leaves"])
def test_column_sampler_matches_oracle(xb, bylevel, bynode, extra):
    import torch
    n, f = 15000, 24
    x, y = make_regression(n, f, seed=73, missing=0.05)
    cp, cv = make_cuts(x, 32)
    page = _page(xb, x, cp, cv)
    gb = orc.bin_dense(x, cp, cv)
    params = dict(extra, eta=0.5)
    up = xb.GPUHistMaker(page)
    up.configure(params)
    rng = np.random.default_rng(74)
    pred = np.zeros(n, dtype=np.float32)
    for it in range(3):
        seeds = [int(v) for v in rng.integers(0, 2 ** 32, size=up.max_nodes + 300)]
        tree_set = np.sort(rng.permutation(f)[:18]).astype(np.int32)
        up.set_column_sampler(tree_set, bylevel, bynode, seeds)
        gp = squared_error_gpair(pred, y)
        tree = up.update(torch.from_numpy(gp).cuda())
        o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), want_hist=True, column_seeds=seeds, tree_set=tree_set,
                            colsample_bylevel=bylevel, colsample_bynode=bynode)
        _compare_tree(xb, up, tree, o, cp, cv, hists=True)
        assert up.last_draw_count() == o["draws_used"]
        pred = orc.update_prediction_cache(o["nodes"], o["position"], pred)
    up.set_column_sampler(None, 1.0, 1.0, None)
    gp = squared_error_gpair(np.zeros(n, np.float32), y)
    t_plain = up.update(torch.from_numpy(gp).cuda())
    _compare_tree(xb, up, t_plain, orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params)), cp, cv, hists=False)


@pytest.mark.parametrize("extra", [{"colsample_bynode": 0.5, "colsample_bylevel": 0.6},
                                   {"colsample_bynode": 0.7, "colsample_bylevel": 0.7, "colsample_bytree": 0.8, "subsample": 0.8,
                                    "grow_policy": "lossguide", "max_leaves": 24, "max_depth": 0}],
                         ids=["bylevel_bynode", "all_three_lossguide"])
def test_colsample_bynode_with_bylevel_through_the_shim_matches_reference(xb, extra):
    """Level and node draws interleave in one engine in the reference; the shim hands the stream of outputs to the
    device sampler and advances the engine by the count it reports."""
    import torch
    rc = _shim_or_skip()
    x, y = make_regression(30000, 40, seed=84)
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": 64, "max_depth": 5, "eta": 0.3,
              "objective": "reg:squarederror", "base_score": 0.5}
    params.update(extra)

    def train(updater):
        dm = rc.DMatrix.from_cuda(torch.from_numpy(x).cuda(), kind="cuda")
        dm.set_label(y)
        bst = rc.Booster(dict(params, updater=updater), dm)
        for i in range(4):
            bst.update(i)
        return _model_trees(bst), bst.predict_margin(dm)
    ref_trees, ref_margin = train("grow_gpu_hist_reference")
    our_trees, our_margin = train("grow_b200_hist")
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))


def test_colsample_bynode_with_interaction_constraints_through_the_shim(xb):
    """Query(the node's own sampled list, nidx) with overlapping sets, through the device column sampler."""
    import torch
    rc = _shim_or_skip()
    x, y = make_regression(30000, 12, seed=85)
    y = (y + x[:, 0] * x[:, 1] + x[:, 2] * x[:, 3] - x[:, 4] * x[:, 5]).astype(np.float32)
    params = {"tree_method": "hist", "device": "cuda:0", "max_bin": 64, "max_depth": 6, "eta": 0.3,
              "objective": "reg:squarederror", "base_score": 0.5, "colsample_bynode": 0.7, "colsample_bylevel": 0.9,
              "interaction_constraints": "[[0, 1, 2], [2, 3, 4], [4, 5, 6], [6, 7, 8, 9, 10, 11]]"}

    def train(updater):
        dm = rc.DMatrix.from_cuda(torch.from_numpy(x).cuda(), kind="cuda")
        dm.set_label(y)
        bst = rc.Booster(dict(params, updater=updater), dm)
        for i in range(4):
            bst.update(i)
        return _model_trees(bst), bst.predict_margin(dm)
    ref_trees, ref_margin = train("grow_gpu_hist_reference")
    our_trees, our_margin = train("grow_b200_hist")
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))
