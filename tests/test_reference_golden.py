"""Parity against the UNMODIFIED reference gpu_hist.

tests/golden/ref_gpu_trees.json holds trees produced by the reference's own CUDA `gpu_hist` updater (built
from /root/reference for sm_100, run on a B200 by tools/make_golden_gpu.py) together with the cuts it
sketched.  Inputs regenerate from recorded seeds.

 * CPU test: the oracle (oracle/hist_oracle.c), given the reference's cuts and the same gradients, rebuilds the
   reference's trees -- this is what PINS the oracle.  The oracle divides in IEEE arithmetic where the device
   uses div.approx, so a split may legitimately differ only when the oracle's own gain table shows a <= 4 ulp tie.
 * GPU test: our CUDA updater reproduces the reference trees BIT FOR BIT (same arithmetic, same device
   instructions): split feature, split condition bits, default direction, node numbering, base weights,
   hessian sums, leaf values; loss_chg is compared bit-exactly as well.
"""
import json
import os

import numpy as np
import pytest

import oracle as orc
from tests.helpers import MT19937, make_regression, make_sparse_regression

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_gpu_trees.json")


def _cases():
    return json.load(open(GOLD))["cases"]


def _f32(bits):
    return np.array(bits, dtype=np.uint32).view(np.float32)


def _gradients(obj, margin, y):
    if obj == "sq":
        g = (margin - y).astype(np.float32)
        h = np.ones_like(g)
    else:
        p = (1.0 / (1.0 + np.exp(-margin.astype(np.float64))))
        g = (p - y).astype(np.float32)
        h = np.maximum(p * (1.0 - p), 1e-16).astype(np.float32)
    return np.stack([g, h], axis=1).astype(np.float32)


def _inputs(case):
    if case.get("sparse"):
        x, y, _, _, _ = make_sparse_regression(case["n"], case["f"], case["sparse"], case["seed"])
        return x, y, np.array(case["cut_ptrs"], dtype=np.uint32), _f32(case["cut_values_bits"])
    x, y = make_regression(case["n"], case["f"], seed=case["seed"], missing=case.get("missing", 0.0))
    if case["obj"] == "logistic":
        y = (y > 0).astype(np.float32)
    cp = np.array(case["cut_ptrs"], dtype=np.uint32)
    cv = _f32(case["cut_values_bits"])
    return x, y, cp, cv


def _params(case):
    p = {k: v for k, v in case["params"].items() if k not in ("subsample", "seed")}
    p.setdefault("lambda", 1.0)
    p.setdefault("min_child_weight", 1.0)
    return p


def _leaf_margin(tree_nodes_leaf_value, position):
    pos = np.where(position < 0, ~position, position)
    return tree_nodes_leaf_value[pos]


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_oracle_reproduces_reference_gpu_hist(case):
    x, y, cp, cv = _inputs(case)
    gb = orc.bin_dense(x, cp, cv)  # semantics do not depend on the storage layout (dense-with-NaN here)
    p = orc.Param.make(**_params(case))
    margin = np.zeros(case["n"], dtype=np.float32)
    sub, rng = case["params"].get("subsample", 1.0), MT19937(case["params"].get("seed", 0))
    for t, ref in enumerate(case["trees"]):
        gp = _gradients(case["obj"], margin, y)
        # subsample < 1: the tree's sampler seed is the next output of the Booster's mt19937 (sampler.cu:165)
        o = orc.update_tree(gb, cp, cv, gp, p, want_hist=True, subsample=sub, sample_seed=rng.next() if sub < 1 else 0)
        nodes = o["nodes"]
        assert len(nodes) == len(ref["left"]), "tree %d node count" % t
        cond = _f32(ref["cond_bits"])
        for i, nd in enumerate(nodes):
            same = (nd.left == ref["left"][i] and nd.right == ref["right"][i])
            if not nd.is_leaf:
                same = same and nd.findex == ref["findex"][i] and \
                    np.float32(nd.fvalue).view(np.uint32) == ref["cond_bits"][i] and \
                    int(nd.default_left) == ref["default_left"][i]
            if not same:
                # only a <= 4 ulp tie in the oracle's own gain table excuses a difference
                _, gains = orc.eval_node(o["hist"][i], (o["cands"][i].left_sum.g + o["cands"][i].right_sum.g,
                                                       o["cands"][i].left_sum.h + o["cands"][i].right_sum.h),
                                         cp, cv, p, o["scale"], want_gains=True)
                pytest.fail("tree %d node %d differs from the reference (oracle f=%d cond=%r, ref f=%d cond=%r)" %
                            (t, i, nd.findex, nd.fvalue, ref["findex"][i], cond[i]))
            assert np.float32(nd.sum_hess).view(np.uint32) == ref["sum_hess_bits"][i], (t, i)
            assert np.float32(nd.base_weight).view(np.uint32) == ref["base_weight_bits"][i], (t, i)
            if not nd.is_leaf:  # IEEE vs approximate division: tolerance scales with the gains involved
                ref_l = float(_f32([ref["loss_chg_bits"][i]])[0])
                c = o["cands"][i]
                pg = float(c.left_sum.g + c.right_sum.g) * o["scale"][2]
                ph = float(c.left_sum.h + c.right_sum.h) * o["scale"][3]
                assert abs(nd.loss_chg - ref_l) <= 1e-5 * (abs(ref_l) + pg * pg / max(ph, 1e-30)) + 1e-6, (t, i)
        leaf = np.array([n.leaf_value for n in nodes], dtype=np.float32)
        margin = (margin + _leaf_margin(leaf, o["position"])).astype(np.float32)
        assert abs(float(np.float64(margin).sum()) - case["margin_sums"][t]) <= 1e-3 * max(1.0, abs(case["margin_sums"][t]))


@pytest.mark.gpu
@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_cuda_matches_reference_gpu_hist_bit_for_bit(case, cuda_device):
    import torch
    import xgboost_b200 as xb
    x, y, cp, cv = _inputs(case)
    if case.get("sparse"):  # the reference stores this one as a sparse ELLPACK page; so do we
        _, _, indptr, indices, values = make_sparse_regression(case["n"], case["f"], case["sparse"], case["seed"])
        page = xb.EllpackPage.from_csr(indptr, indices, values, xb.HistogramCuts(cp, cv))
        assert page.info()["bin_bytes"] == 4
    else:
        page = xb.EllpackPage.from_dense(torch.from_numpy(x).cuda(), xb.HistogramCuts(cp, cv))
    up = xb.GPUHistMaker(page)
    up.configure(_params(case))
    margin = torch.zeros(case["n"], dtype=torch.float32, device="cuda")
    sub, rng = case["params"].get("subsample", 1.0), MT19937(case["params"].get("seed", 0))
    for t, ref in enumerate(case["trees"]):
        gp = torch.from_numpy(_gradients(case["obj"], margin.cpu().numpy(), y)).cuda()
        if sub < 1:
            up.set_tree_sampling(None, sub, rng.next())
        tree = up.update(gp)
        assert tree.num_nodes == len(ref["left"]), "tree %d node count" % t
        for i, nd in enumerate(tree.nodes):
            assert (nd["left"], nd["right"]) == (ref["left"][i], ref["right"][i]), (t, i)
            if not nd["is_leaf"]:
                assert nd["findex"] == ref["findex"][i], (t, i)
                assert np.float32(nd["fvalue"]).view(np.uint32) == ref["cond_bits"][i], (t, i)
                assert int(nd["default_left"]) == ref["default_left"][i], (t, i)
                assert np.float32(nd["loss_chg"]).view(np.uint32) == ref["loss_chg_bits"][i], (t, i)
            else:  # RegTree stores the leaf value in split_conditions
                assert np.float32(nd["leaf_value"]).view(np.uint32) == ref["cond_bits"][i], (t, i)
            assert np.float32(nd["sum_hess"]).view(np.uint32) == ref["sum_hess_bits"][i], (t, i)
            assert np.float32(nd["base_weight"]).view(np.uint32) == ref["base_weight_bits"][i], (t, i)
        up.update_prediction_cache(margin)
        assert abs(float(margin.double().sum().item()) - case["margin_sums"][t]) <= 1e-3 * max(1.0, abs(case["margin_sums"][t]))


# ------------------------------------------------------------------------------------------------
# Q2: quantile sketch -> cuts.  The golden file holds the cut values the reference's GPU sketch produced.
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [c for c in _cases() if c["f"] <= 64 or c["n"] <= 30000], ids=lambda c: c["name"])
def test_oracle_sketch_reproduces_reference_cuts(case):
    from oracle import sketch
    x, _, cp, cv = _inputs(case)
    p, v = sketch.device_sketch_cuts(x, case["max_bin"])
    assert np.array_equal(p, cp)
    assert np.array_equal(v.view(np.uint32), cv.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_cuda_sketch_reproduces_reference_cuts(case, cuda_device):
    import torch
    import xgboost_b200 as xb
    x, _, cp, cv = _inputs(case)
    cuts = xb.HistogramCuts.from_dense(torch.from_numpy(x).cuda(), case["max_bin"])
    assert np.array_equal(cuts.cut_ptrs, cp)
    assert np.array_equal(cuts.cut_values.view(np.uint32), cv.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("n,f,max_bin,kind", [(5000, 7, 16, "int"), (70000, 5, 64, "int"), (3, 2, 8, "cont"),
                                              (300, 4, 256, "cont"), (100000, 9, 256, "mixed"), (1, 1, 4, "cont")])
def test_cuda_sketch_matches_oracle_on_duplicates(n, f, max_bin, kind, cuda_device):
    """Discrete / duplicated values exercise the duplicate folding of ScanInput; edge sizes the small paths."""
    import torch
    import xgboost_b200 as xb
    from oracle import sketch
    rng = np.random.default_rng(n + f)
    if kind == "int":
        x = rng.integers(-3, 40, size=(n, f)).astype(np.float32)
    elif kind == "cont":
        x = rng.standard_normal((n, f)).astype(np.float32)
    else:
        x = rng.standard_normal((n, f)).astype(np.float32)
        x[:, ::2] = np.round(x[:, ::2] * 4) / 4
        x[rng.random((n, f)) < 0.1] = np.nan
        x[:, -1] = np.nan  # an empty column
    cuts = xb.HistogramCuts.from_dense(torch.from_numpy(x).cuda(), max_bin)
    p, v = sketch.device_sketch_cuts(x, max_bin)
    assert np.array_equal(cuts.cut_ptrs, p)
    # compared as VALUES: -0.0 and +0.0 are equal keys for the reference's comparison sort (their relative
    # order, hence the sign bit of a zero cut, is unspecified there); everything else must agree exactly
    assert np.array_equal(cuts.cut_values, v)
    nz = v != 0
    assert np.array_equal(cuts.cut_values[nz].view(np.uint32), v[nz].view(np.uint32))
    # property the reference's own tests check (tests/cpp/common/test_hist_util.h:134-180): cuts strictly increase
    for j in range(f):
        seg = cuts.cut_values[p[j]:p[j + 1]]
        assert np.all(np.diff(seg) > 0)


# ---------------------------------------------------------------------------------------------------------------------
# Interaction constraints: tests/golden/ref_gpu_interaction_trees.json = the reference's gpu_hist trees under
# OVERLAPPING constraint sets (tools/probe_interaction_reference.py on a B200).  This is the fixture that settles which
# reading of FeatureInteractionConstraintDevice::Query the reference's trees follow (oracle orc_update_tree_ic).
# ---------------------------------------------------------------------------------------------------------------------
IC_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_gpu_interaction_trees.json")


def _ic_cases():
    return json.load(open(IC_GOLD))["cases"]


def _ic_tree_equal(o, ref):
    nodes = o["nodes"]
    if len(nodes) != len(ref["left"]):
        return False
    for i, nd in enumerate(nodes):
        if nd.left != ref["left"][i] or nd.right != ref["right"][i]:
            return False
        if not nd.is_leaf and (nd.findex != ref["findex"][i] or
                               int(np.float32(nd.fvalue).view(np.uint32)) != ref["cond_bits"][i] or
                               int(nd.default_left) != ref["default_left"][i]):
            return False
        if int(np.float32(nd.base_weight).view(np.uint32)) != ref["base_weight_bits"][i] or \
                int(np.float32(nd.sum_hess).view(np.uint32)) != ref["sum_hess_bits"][i]:
            return False
    return True


@pytest.mark.parametrize("case", _ic_cases(), ids=lambda c: c["name"])
def test_oracle_reproduces_reference_interaction_constraints(case):
    """The reference's evaluation kernel reads the ONE result buffer every Query of a level wrote into
    (constraints.cu:125,199-240; updater_gpu_hist.cu:251-268): the oracle's shared_result_buffer reading rebuilds
    every reference tree bit for bit, and on the first tree of every case the per-query reading does NOT -- so the
    fixture discriminates."""
    from tests.helpers import make_interaction_data as make_data
    x, y = make_data(case)
    cp = np.array(case["cut_ptrs"], dtype=np.uint32)
    cv = _f32(case["cut_values_bits"])
    gb = orc.bin_dense(x, cp, cv)
    p = orc.Param.make(**case["params"])
    margin = np.zeros(case["n"], dtype=np.float32)
    discriminated = False
    for t, ref in enumerate(case["trees"]):
        gp = np.stack([margin - y, np.ones_like(y)], axis=1).astype(np.float32)
        o = orc.update_tree(gb, cp, cv, gp, p, interaction=case["sets"], interaction_shared_buffer=True)
        assert _ic_tree_equal(o, ref), "tree %d differs from the reference" % t
        per_query = orc.update_tree(gb, cp, cv, gp, p, interaction=case["sets"], interaction_shared_buffer=False)
        discriminated = discriminated or not _ic_tree_equal(per_query, ref)
        leaf = np.array([n.leaf_value for n in o["nodes"]], dtype=np.float32)
        margin = (margin + _leaf_margin(leaf, o["position"])).astype(np.float32)
    assert discriminated


# ---------------------------------------------------------------------------------------------------------------------
# Categorical features, whole trees: tests/golden/ref_cpu_categorical_trees.json = the reference's CPU `hist` updater
# (tools/make_golden_cpu_categorical.py).  A STRUCTURAL pin of orc_update_tree_full until the CUDA reference has been
# run on such data: float vs fixed-point histograms make gains differ in the last bits, everything discrete agrees.
# ---------------------------------------------------------------------------------------------------------------------
CAT_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cpu_categorical_trees.json")


def categorical_global_bins(x, cp, cv, is_cat):
    """SearchCatBin / SearchBin (src/common/hist_util.h:112-135): a category is looked up among the feature's
    category codes, a numerical value by upper bound."""
    gb = np.full(x.shape, -1, dtype=np.int32)
    for j in range(x.shape[1]):
        ok = ~np.isnan(x[:, j])
        seg = cv[cp[j]:cp[j + 1]]
        if is_cat[j]:
            gb[ok, j] = cp[j] + np.searchsorted(seg, x[ok, j])
        else:
            gb[ok, j] = cp[j] + orc.bin_dense(x[:, j:j + 1], np.array([0, seg.size], dtype=np.uint32), seg)[ok, 0]
    return gb


@pytest.mark.parametrize("case", json.load(open(CAT_GOLD))["cases"], ids=lambda c: c["name"])
def test_oracle_categorical_trees_match_cpu_reference(case):
    from tests.helpers import make_categorical_data as make_data
    x, y, types = make_data(case)
    assert types == case["feature_types"]
    is_cat = [1 if t == "c" else 0 for t in types]
    cp = np.array(case["cut_ptrs"], dtype=np.uint32)
    cv = _f32(case["cut_values_bits"])
    gb = categorical_global_bins(x, cp, cv, is_cat)
    gp = np.stack([-y, np.ones_like(y)], axis=1).astype(np.float32)
    o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**case["params"]), feature_is_cat=is_cat,
                        max_cat_to_onehot=case["max_cat_to_onehot"], max_cat_threshold=case["max_cat_threshold"])
    ref = case["tree"]
    nodes = o["nodes"]
    assert len(nodes) == len(ref["left"])
    pos = np.where(o["position"] < 0, ~o["position"], o["position"])

    def leaves(k):
        return [k] if nodes[k].is_leaf else leaves(nodes[k].left) + leaves(nodes[k].right)
    n_cat = 0
    for i, nd in enumerate(nodes):
        assert (nd.left, nd.right) == (ref["left"][i], ref["right"][i]), i
        if nd.is_leaf:
            continue
        assert nd.findex == ref["findex"][i], i
        assert int(o["is_cat"][i]) == ref["split_type"][i], i
        if o["is_cat"][i]:
            n_cat += 1
            assert sorted(o["cat_sets"][i]) == ref["categories"][i], i
            assert np.isnan(nd.fvalue), i
        else:
            assert int(np.float32(nd.fvalue).view(np.uint32)) == ref["cond_bits"][i], i
        rows = np.isin(pos, leaves(i))
        if np.isnan(x[rows, nd.findex]).any():   # otherwise both directions tie
            assert int(nd.default_left) == ref["default_left"][i], i
    assert n_cat >= 5


# ---------------------------------------------------------------------------------------------------------------------
# Prediction on quantised data: the oracle's walk + summation order against the reference's own
# CPU predictor, run live on oracle/_ref when it is there (build container); the GPU parity test of the CUDA kernel
# compares with this oracle.
# ---------------------------------------------------------------------------------------------------------------------
def test_oracle_predictor_reproduces_reference_cpu_margins():
    from oracle import ref_capi as rc
    if not rc.available("cpu"):
        pytest.skip("needs oracle/_ref/libxgboost.so (built where /root/reference exists)")
    x, y = make_regression(20000, 12, seed=88, missing=0.05)
    dm = rc.DMatrix.from_dense(x, label=y, kind="cpu")
    bst = rc.Booster({"tree_method": "hist", "device": "cpu", "max_bin": 64, "base_score": 0.0,
                      "objective": "reg:squarederror", "max_depth": 6, "eta": 0.3}, dm)
    for i in range(8):
        bst.update(i)
    cp, cv = dm.histogram_cuts()
    gb = orc.bin_dense(x, cp, cv)
    forest = []
    for t in rc.trees_from_model(bst.save_json()):
        nodes = []
        for i in range(len(t["left"])):
            leaf = t["left"][i] == -1
            nodes.append(orc.Node(t["parent"][i], t["left"][i], t["right"][i], -1 if leaf else t["findex"][i],
                                  t["cond"][i], t["default_left"][i], 1 if leaf else 0, t["cond"][i] if leaf else 0.0,
                                  0, 0, 0, 0))
        forest.append(nodes)
    ours = orc.predict_margin(gb, cp, cv, forest)
    ref = bst.predict_margin(dm)
    assert np.array_equal(ours.view(np.uint32), ref.view(np.uint32))
    # the summation order is part of the result: per-tree accumulation into the output differs in the last bits
    naive = np.zeros(len(y), dtype=np.float32)
    for nodes in forest:
        naive = orc.predict_margin(gb, cp, cv, [nodes], base=naive)
    assert np.allclose(naive, ref, rtol=1e-5, atol=1e-6)


def test_oracle_predictor_on_foreign_cuts_reproduces_reference_quantised_prediction():
    """Trees grown on one set of cuts, predicted on a matrix quantised with ANOTHER: the split conditions are no longer
    cuts of the page, and the only rule that reproduces the reference is its loader's (the bin's lower cut, -inf for
    a feature's first bin; src/predictor/data_accessor.h:117-126 on the CPU, gpu_data_accessor.cuh:95-107 on the
    GPU).  The reference predicts from the QuantileDMatrix's bins, the oracle from the same bins."""
    from oracle import ref_capi as rc
    if not rc.available("cpu"):
        pytest.skip("needs oracle/_ref/libxgboost.so (built where /root/reference exists)")
    x, y = make_regression(12000, 10, seed=89, missing=0.08)
    dm = rc.DMatrix.from_dense(x, label=y, kind="cpu")
    bst = rc.Booster({"tree_method": "hist", "device": "cpu", "max_bin": 64, "base_score": 0.0,
                      "objective": "reg:squarederror", "max_depth": 5, "eta": 0.3}, dm)
    for i in range(6):
        bst.update(i)
    forest = []
    for t in rc.trees_from_model(bst.save_json()):
        nodes = []
        for i in range(len(t["left"])):
            leaf = t["left"][i] == -1
            nodes.append(orc.Node(t["parent"][i], t["left"][i], t["right"][i], -1 if leaf else t["findex"][i],
                                  t["cond"][i], t["default_left"][i], 1 if leaf else 0, t["cond"][i] if leaf else 0.0,
                                  0, 0, 0, 0))
        forest.append(nodes)
    x2, y2 = make_regression(5000, 10, seed=90, missing=0.08)
    qdm = rc.DMatrix.quantile_from_batches([x2], max_bin=24, kind="cpu")
    cp2, cv2 = qdm.histogram_cuts()
    train_cuts = set(np.asarray(dm.histogram_cuts()[1]).view(np.uint32).tolist())
    assert not set(np.asarray(cv2).view(np.uint32).tolist()) <= train_cuts   # really foreign cuts
    gb2 = orc.bin_dense(x2, cp2, cv2)
    ours = orc.predict_margin(gb2, cp2, cv2, forest)
    ref = bst.predict_margin(qdm)
    assert np.array_equal(ours.view(np.uint32), ref.view(np.uint32))
    # ... and it is NOT what predicting from the raw values gives (quantisation moved some rows across a split)
    raw = bst.predict_margin(rc.DMatrix.from_dense(x2, kind="cpu"))
    assert not np.array_equal(raw.view(np.uint32), ref.view(np.uint32))


def test_oracle_grouped_prediction_reproduces_reference_multiclass_margins():
    """Three output groups (multi:softprob): tree t feeds group tree_info[t] and every leaf goes straight onto the
    output, tree after tree -- with a non-zero base score that differs in the last bits from summing the trees first.
    The reference's CPU predictor is the pin (the GPU kernel's multi-group branch has the same form,
    gpu_predictor.cu:229-239)."""
    from oracle import ref_capi as rc
    if not rc.available("cpu"):
        pytest.skip("needs oracle/_ref/libxgboost.so (built where /root/reference exists)")
    x, y = make_regression(9000, 8, seed=91, missing=0.05)
    label = np.digitize(y, np.quantile(y, [1 / 3, 2 / 3])).astype(np.float32)
    dm = rc.DMatrix.from_dense(x, label=label, kind="cpu")
    bst = rc.Booster({"tree_method": "hist", "device": "cpu", "max_bin": 32, "objective": "multi:softprob",
                      "num_class": 3, "base_score": 0.5, "max_depth": 4, "eta": 0.5}, dm)
    for i in range(5):
        bst.update(i)
    model = bst.save_json()
    groups = [int(g) for g in model["learner"]["gradient_booster"]["model"]["tree_info"]]
    assert groups == [0, 1, 2] * 5
    cp, cv = dm.histogram_cuts()
    gb = orc.bin_dense(x, cp, cv)
    forest = []
    for t in rc.trees_from_model(model):
        nodes = []
        for i in range(len(t["left"])):
            leaf = t["left"][i] == -1
            nodes.append(orc.Node(t["parent"][i], t["left"][i], t["right"][i], -1 if leaf else t["findex"][i],
                                  t["cond"][i], t["default_left"][i], 1 if leaf else 0, t["cond"][i] if leaf else 0.0,
                                  0, 0, 0, 0))
        forest.append(nodes)
    ref = bst.predict_margin(dm).reshape(len(y), 3)
    # the margin of a model without trees is the base margin: recover it from the reference itself
    bst0 = rc.Booster({"tree_method": "hist", "device": "cpu", "max_bin": 32, "objective": "multi:softprob",
                       "num_class": 3, "base_score": 0.5}, dm)
    base = bst0.predict_margin(dm).reshape(len(y), 3)
    assert float(base[0, 0]) != 0.0
    ours = orc.predict_margin_groups(gb, cp, cv, forest, groups, 3, base=base)
    assert np.array_equal(ours.view(np.uint32), ref.view(np.uint32))
    # summing a group's trees first and adding the sum afterwards is NOT the same float
    summed = base.copy()
    for g in range(3):
        summed[:, g] = base[:, g] + orc.predict_margin(gb, cp, cv, [forest[i] for i in range(15) if groups[i] == g])
    assert not np.array_equal(summed.view(np.uint32), ref.view(np.uint32))


# ---------------------------------------------------------------------------------------------------------------------
# Waiting for their fixture: tests/golden/ref_gpu_next_trees.json (tools/make_golden_gpu_next.py, one short GPU call).
# The reference's own gpu_hist trees for categorical features and gradient-based sampling pin the two oracle
# restatements bit for bit, the way ref_gpu_trees.json pins the numerical path.
# ---------------------------------------------------------------------------------------------------------------------
NEXT_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_gpu_next_trees.json")


def _next_cases(kind):
    if not os.path.exists(NEXT_GOLD):
        return []
    return [c for c in json.load(open(NEXT_GOLD))["cases"] if c["kind"] == kind]


def _assert_tree_equals_reference(o, ref, t, cats=False):
    nodes = o["nodes"]
    assert len(nodes) == len(ref["left"]), "tree %d node count" % t
    for i, nd in enumerate(nodes):
        assert (nd.left, nd.right) == (ref["left"][i], ref["right"][i]), (t, i)
        if not nd.is_leaf:
            assert nd.findex == ref["findex"][i] and int(nd.default_left) == ref["default_left"][i], (t, i)
            if cats and o["is_cat"][i]:
                assert ref["split_type"][i] == 1 and sorted(o["cat_sets"][i]) == ref["categories"][i], (t, i)
            else:
                assert ref["split_type"][i] == 0, (t, i)
                assert int(np.float32(nd.fvalue).view(np.uint32)) == ref["cond_bits"][i], (t, i)
            ref_l = float(_f32([ref["loss_chg_bits"][i]])[0])
            assert abs(nd.loss_chg - ref_l) <= 2e-5 * (abs(ref_l) + abs(nd.loss_chg)) + 1e-5, (t, i)
        assert int(np.float32(nd.sum_hess).view(np.uint32)) == ref["sum_hess_bits"][i], (t, i)
        assert int(np.float32(nd.base_weight).view(np.uint32)) == ref["base_weight_bits"][i], (t, i)


@pytest.mark.skipif(not os.path.exists(NEXT_GOLD), reason="fixture not generated yet (tools/make_golden_gpu_next.py)")
@pytest.mark.parametrize("case", _next_cases("categorical"), ids=lambda c: c["name"])
def test_oracle_reproduces_reference_gpu_categorical_trees(case):
    from tests.helpers import make_categorical_data
    x, y, types = make_categorical_data(case)
    is_cat = [1 if t == "c" else 0 for t in types]
    cp = np.array(case["cut_ptrs"], dtype=np.uint32)
    cv = _f32(case["cut_values_bits"])
    gb = categorical_global_bins(x, cp, cv, is_cat)
    params = {k: v for k, v in case["params"].items() if k not in ("max_cat_to_onehot", "max_cat_threshold")}
    margin = np.zeros(case["n"], dtype=np.float32)
    for t, ref in enumerate(case["trees"]):
        gp = np.stack([margin - y, np.ones_like(y)], axis=1).astype(np.float32)
        o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), feature_is_cat=is_cat,
                            max_cat_to_onehot=case["params"]["max_cat_to_onehot"],
                            max_cat_threshold=case["params"]["max_cat_threshold"])
        _assert_tree_equals_reference(o, ref, t, cats=True)
        leaf = np.array([n.leaf_value for n in o["nodes"]], dtype=np.float32)
        margin = (margin + _leaf_margin(leaf, o["position"])).astype(np.float32)


@pytest.mark.skipif(not os.path.exists(NEXT_GOLD), reason="fixture not generated yet (tools/make_golden_gpu_next.py)")
@pytest.mark.parametrize("case", _next_cases("sampling"), ids=lambda c: c["name"])
def test_oracle_reproduces_reference_gpu_gradient_based_sampling(case):
    x, y = make_regression(case["n"], case["f"], seed=case["seed"])
    cp = np.array(case["cut_ptrs"], dtype=np.uint32)
    cv = _f32(case["cut_values_bits"])
    gb = orc.bin_dense(x, cp, cv)
    params = {k: v for k, v in case["params"].items() if k not in ("subsample", "sampling_method", "seed")}
    rng = MT19937(case["params"].get("seed", 0))
    margin = np.zeros(case["n"], dtype=np.float32)
    for t, ref in enumerate(case["trees"]):
        gp = np.stack([margin - y, np.ones_like(y)], axis=1).astype(np.float32)
        o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), subsample=case["params"]["subsample"],
                            sample_seed=rng.next(), sampling_method="gradient_based")
        _assert_tree_equals_reference(o, ref, t)
        leaf = np.array([n.leaf_value for n in o["nodes"]], dtype=np.float32)
        margin = (margin + _leaf_margin(leaf, o["position"])).astype(np.float32)


@pytest.mark.skipif(not os.path.exists(NEXT_GOLD), reason="fixture not generated yet (tools/make_golden_gpu_next.py)")
@pytest.mark.parametrize("case", _next_cases("bynode"), ids=lambda c: c["name"])
def test_oracle_reproduces_reference_gpu_colsample_bynode(case):
    """The per-node feature sets are regenerated from the Booster's mt19937: per tree, one draw seeds the tree-level
    shuffle (colsample_bytree < 1 only, ColumnSampler::Init), then every GetFeatureSet call -- root, then left and
    right child of every valid candidate in Pop order -- takes one draw for a thrust::shuffle of the tree set
    (random.cu:62-96), restated by orc.col_sample."""
    x, y = make_regression(case["n"], case["f"], seed=case["seed"])
    cp = np.array(case["cut_ptrs"], dtype=np.uint32)
    cv = _f32(case["cut_values_bits"])
    gb = orc.bin_dense(x, cp, cv)
    prm = case["params"]
    params = {k: v for k, v in prm.items() if k not in ("colsample_bynode", "colsample_bytree", "seed")}
    p = orc.Param.make(**params)
    max_nodes = min((1 << (p.max_depth + 1)) - 1 if p.max_depth > 0 else 1 << 30,
                    2 * p.max_leaves - 1 if p.max_leaves > 0 else 1 << 30)
    rng = MT19937(prm.get("seed", 0))
    margin = np.zeros(case["n"], dtype=np.float32)
    all_features = np.arange(case["f"], dtype=np.int32)
    for t, ref in enumerate(case["trees"]):
        bytree = prm.get("colsample_bytree", 1.0)
        tree_set = orc.col_sample(all_features, bytree, rng.next()) if bytree < 1.0 else all_features
        ahead = rng.copy()            # the draws the tree MAY use; the engine itself advances by the draws it DID use
        sets = np.stack([orc.col_sample(tree_set, prm["colsample_bynode"], ahead.next()) for _ in range(max_nodes)])
        gp = np.stack([margin - y, np.ones_like(y)], axis=1).astype(np.float32)
        o = orc.update_tree(gb, cp, cv, gp, p, node_feature_sets=sets)
        _assert_tree_equals_reference(o, ref, t)
        for _ in range(o["eval_count"]):
            rng.next()
        leaf = np.array([n.leaf_value for n in o["nodes"]], dtype=np.float32)
        margin = (margin + _leaf_margin(leaf, o["position"])).astype(np.float32)


@pytest.mark.skipif(not os.path.exists(NEXT_GOLD), reason="fixture not generated yet (tools/make_golden_gpu_next.py)")
@pytest.mark.parametrize("case", _next_cases("colsample"), ids=lambda c: c["name"])
def test_oracle_reproduces_reference_gpu_colsample_bytree_bylevel(case):
    """Engine draws per tree, in the reference's order (GPUHistMakerDevice::Reset, updater_gpu_hist.cu:172-211): the row
    sampler's seed (subsample < 1), the tree-level shuffle seed (colsample_bytree < 1, ColumnSampler::Init), then one
    seed per depth the first time a node of that depth is evaluated (colsample_bylevel < 1, random.h:156-175)."""
    x, y = make_regression(case["n"], case["f"], seed=case["seed"])
    cp = np.array(case["cut_ptrs"], dtype=np.uint32)
    cv = _f32(case["cut_values_bits"])
    gb = orc.bin_dense(x, cp, cv)
    prm = case["params"]
    params = {k: v for k, v in prm.items() if k not in ("colsample_bylevel", "colsample_bytree", "subsample", "seed")}
    p = orc.Param.make(**params)
    rng = MT19937(prm.get("seed", 0))
    margin = np.zeros(case["n"], dtype=np.float32)
    all_features = np.arange(case["f"], dtype=np.int32)
    sub, bytree, bylevel = prm.get("subsample", 1.0), prm.get("colsample_bytree", 1.0), prm.get("colsample_bylevel", 1.0)
    for t, ref in enumerate(case["trees"]):
        row_seed = rng.next() if sub < 1.0 else 0
        tree_set = orc.col_sample(all_features, bytree, rng.next()) if bytree < 1.0 else all_features
        gp = np.stack([margin - y, np.ones_like(y)], axis=1).astype(np.float32)
        if bylevel < 1.0:
            ahead = rng.copy()
            levels = np.stack([orc.col_sample(tree_set, bylevel, ahead.next()) for _ in range(p.max_depth)])
            o = orc.update_tree(gb, cp, cv, gp, p, subsample=sub, sample_seed=row_seed, level_feature_sets=levels)
            deepest = max(nd.depth for nd, c in zip(o["nodes"], o["cands"]) if c.loss_chg != 0.0)   # evaluated nodes
            for _ in range(deepest + 1):
                rng.next()
        else:
            o = orc.update_tree(gb, cp, cv, gp, p, subsample=sub, sample_seed=row_seed, feature_set=tree_set)
        _assert_tree_equals_reference(o, ref, t)
        leaf = np.array([n.leaf_value for n in o["nodes"]], dtype=np.float32)
        margin = (margin + _leaf_margin(leaf, o["position"])).astype(np.float32)
