"""CPU tests: pin the oracle (oracle/hist_oracle.c) against the reference's own known answers.
These run without a GPU."""
import json
import os

import numpy as np
import pytest

import oracle as orc


def _dummy_scale():
    # DummyRoundingFactor (test_evaluate_splits.cu:21-27): quantiser built from one pair {1000, 1000}
    scale, _ = orc.quantiser(np.array([[1000.0, 1000.0]], dtype=np.float32))
    return scale


def _to_fixed(pairs, scale, as_float32=True):
    a = np.array(pairs, dtype=np.float64).reshape(-1, 2)
    if as_float32:  # ConvertToInteger goes through GradientPair (float)
        return orc.quantise(a.astype(np.float32), scale)
    out = np.empty((a.shape[0], 2), dtype=np.int64)
    out[:, 0] = (a[:, 0] * scale[0]).astype(np.int64)
    out[:, 1] = (a[:, 1] * scale[1]).astype(np.int64)
    return out


def test_dummy_quantiser_scale():
    scale = _dummy_scale()
    # max=1000 -> delta just above 1000 -> M = 1024; to_float = 1024 / 2^62 = 2^-52
    assert scale[2] == 2.0 ** -52 and scale[3] == 2.0 ** -52
    assert scale[0] == 2.0 ** 52


def test_quantiser_round_trip():
    # test_histogram.cu:341-355: 16 samples of {1, 1} survive the round trip exactly
    g = np.ones((16, 2), dtype=np.float32)
    scale, sums = orc.quantiser(g)
    assert sums[0] == 16.0 and sums[2] == 0.0
    q = orc.quantise(g, scale)
    assert np.all(q[:, 0] * scale[2] == 1.0) and np.all(q[:, 1] * scale[3] == 1.0)


def test_rounding_factor_power_of_two_edge():
    # frexp semantics: an exact power of two is bumped to the next one (deterministic.cuh:24-38)
    scale, _ = orc.quantiser(np.array([[1024.0, 0.5]], dtype=np.float32))
    assert scale[2] * 2.0 ** 62 == 2048.0
    scale0, _ = orc.quantiser(np.zeros((4, 2), dtype=np.float32))
    assert scale0[2] * 2.0 ** 62 == 1.0  # frexp(0) -> exp 0


def test_evaluate_known_answers(golden_dir):
    spec = json.load(open(os.path.join(golden_dir, "eval_known_answers.json")))
    scale = _dummy_scale()
    p = orc.Param.make(**spec["param"])
    for case in spec["cases"]:
        hist = _to_fixed(case["hist"], scale)
        parent = _to_fixed([case["parent"]], scale, as_float32=False)[0]
        c = orc.eval_node(hist, parent, case["cut_ptrs"], case["cut_values"], p, scale,
                          feature_set=case["feature_set"])
        e = case["expect"]
        assert c.findex == e["findex"], case["name"]
        if "fvalue" in e:
            assert c.fvalue == e["fvalue"], case["name"]
        if "dir" in e:
            assert c.dir == e["dir"], case["name"]
        if e.get("sum_is_parent"):
            assert (c.left_sum.g + c.right_sum.g, c.left_sum.h + c.right_sum.h) == tuple(parent)
        if "left_sum" in e:
            ls = _to_fixed([e["left_sum"]], scale, as_float32=False)[0]
            rs = _to_fixed([e["right_sum"]], scale, as_float32=False)[0]
            assert (c.left_sum.g, c.left_sum.h) == tuple(ls), case["name"]
            assert (c.right_sum.g, c.right_sum.h) == tuple(rs), case["name"]
        if e.get("loss_chg_negative"):
            assert c.loss_chg < 0.0


def test_add64_as32_known_answers():
    # test_histogram.cu:286-340: mixed signs, values beyond 32 bits in both directions
    acc = 0
    for i in range(1024):
        acc = orc.lib().orc_add64_as32(acc, i if i % 2 == 0 else -i)
    assert acc == sum(i if i % 2 == 0 else -i for i in range(1024))
    acc = 0
    for _ in range(1024):
        acc = orc.lib().orc_add64_as32(acc, (2 ** 32 - 1) // 2)
    assert acc == 1024 * ((2 ** 32 - 1) // 2) and acc > 2 ** 32 - 1
    acc = 0
    for _ in range(1024):
        acc = orc.lib().orc_add64_as32(acc, -(2 ** 31) // 2)
    assert acc == 1024 * (-(2 ** 31) // 2) and acc < -(2 ** 31)


# ---- RandomDataGenerator(rows, cols, sparsity).Seed(seed) restated (tests/cpp/helpers.cc:221-224,354-397)
class _LCG:
    def __init__(self, seed):
        self.mod = (1 << 32) - 1
        self.state = seed % self.mod

    def __call__(self):
        self.state = (61 * self.state + (3 if self.state == 0 else 0)) % self.mod
        return self.state


def _uniform(lcg):
    # SimpleRealUniformDistribution<float>(0, 1): one draw, float(draw) / float(2^32)
    return max(np.float32(np.float32(lcg()) / np.float32(4294967296.0)), np.float32(0.0))


def _generate_csr(rows, cols, sparsity, seed):
    lcg = _LCG(seed)
    out = np.full((rows, cols), np.nan, dtype=np.float32)
    for i in range(rows):
        for j in range(cols):
            g = _uniform(lcg)
            if g >= np.float32(sparsity):
                out[i, j] = _uniform(lcg)
    return out


def test_ellpack_dense_known_answer(golden_dir):
    spec = json.load(open(os.path.join(golden_dir, "ellpack_known_answers.json")))
    x = _generate_csr(spec["n_rows"], spec["n_cols"], 0.0, spec["seed"])
    assert not np.isnan(x).any()
    gb = orc.bin_dense(x, spec["cut_ptrs"], spec["cut_values"])
    want = np.array(spec["dense_global_bins"], dtype=np.int32).reshape(spec["n_rows"], spec["n_cols"])
    assert np.array_equal(gb, want)
    # symbols are feature-local; NumSymbols == 3; pack / read round trip through the reference bit layout
    local = gb - np.array(spec["cut_ptrs"][:-1], dtype=np.int32)[None, :]
    assert local.max() + 1 == spec["dense_n_symbols"]
    buf = orc.ellpack_pack(local.astype(np.uint32), spec["dense_n_symbols"])
    assert orc.lib().orc_symbol_bits(3) == 2
    assert np.array_equal(orc.ellpack_read(buf, 3, local.size), local.ravel().astype(np.uint32))
    # big-endian bit order: first four 2-bit symbols of row 0 live in byte 8
    assert buf[8] == (local[0, 0] << 6 | local[0, 1] << 4 | local[0, 2] << 2 | local[0, 3])


def test_ellpack_sparse_known_answer(golden_dir):
    spec = json.load(open(os.path.join(golden_dir, "ellpack_known_answers.json")))
    x = _generate_csr(spec["n_rows"], spec["n_cols"], spec["sparse_sparsity"], spec["seed"])
    gb = orc.bin_dense(x, spec["cut_ptrs"], spec["cut_values"])
    stride, null = spec["sparse_row_stride"], spec["sparse_null"]
    ell = np.full((spec["n_rows"], stride), null, dtype=np.int64)
    for i in range(spec["n_rows"]):
        present = gb[i][gb[i] >= 0]
        assert present.size <= stride
        ell[i, : present.size] = present
    assert (np.sum(gb >= 0, axis=1).max()) == stride
    assert np.array_equal(ell.ravel(), np.array(spec["sparse_symbols"]))
    assert orc.lib().orc_symbol_bits(25) == 5
    assert orc.lib().orc_ellpack_buffer_size(0, 1) == 9  # empty page: kPadding + 1 (test_ellpack_page.cu:30-34)


def test_partition_semantics():
    # test_row_partitioner.cu:60-108: left rows first, counts right
    cut_ptrs = np.array([0, 2], dtype=np.uint32)
    cut_values = np.array([0.5, 2.0], dtype=np.float32)
    x = np.array([[0.0], [1.0], [0.0], [1.0], [0.0], [1.0]], dtype=np.float32)
    gb = orc.bin_dense(x, cut_ptrs, cut_values)
    out, nl = orc.partition(gb, cut_ptrs, cut_values, np.arange(6), 0, 0.5, True)
    assert nl == 3 and list(out) == [0, 2, 4, 1, 3, 5]
    out, nl = orc.partition(gb, cut_ptrs, cut_values, np.array([3, 4, 5]), 0, -np.inf, True)
    assert nl == 0 and list(out) == [3, 4, 5]


def test_tree_basic_invariants():
    from tests.helpers import make_cuts, make_regression, squared_error_gpair
    x, y = make_regression(2000, 6, seed=1)
    cp, cv = make_cuts(x, 16)
    gb = orc.bin_dense(x, cp, cv)
    gpair = squared_error_gpair(np.zeros_like(y), y)
    p = orc.Param.make(max_depth=4)
    t = orc.update_tree(gb, cp, cv, gpair, p, want_hist=True)
    nodes = t["nodes"]
    assert nodes[0].parent == -1 and len(nodes) % 2 == 1
    gq = orc.quantise(gpair, t["scale"])
    pos = t["position"]
    assert pos.min() >= 0  # hessian is 1 everywhere
    for i, nd in enumerate(nodes):
        if nd.is_leaf:
            continue
        # children hessians add up; histogram of a node sums to its gradient sum per feature
        assert abs(nodes[nd.left].sum_hess + nodes[nd.right].sum_hess - nd.sum_hess) <= 1e-3 * nd.sum_hess
        assert nd.depth < 4 and nd.loss_chg > 1e-6
    # every leaf's rows agree with a manual traversal
    for r in range(0, 2000, 97):
        n = 0
        while not nodes[n].is_leaf:
            b = gb[r, nodes[n].findex]
            n = nodes[n].left if cv[b] <= nodes[n].fvalue else nodes[n].right
        assert pos[r] == n
    # root histogram: every feature's bins sum to the root sum
    h0 = t["hist"][0]
    for f in range(6):
        assert tuple(h0[cp[f]:cp[f + 1]].sum(axis=0)) == tuple(gq.sum(axis=0))


def test_random_weight_restates_thrust_minstd_discard_uniform():
    """RandomWeight(seed)(i) (sampler.cu:36-47): minstd_rand seeded with seed, discard(i), one float draw.
    Checked against a step-by-step big-integer evaluation and against known values of minstd_rand."""
    import oracle as orc
    m, a = 2147483647, 48271

    def ref(seed, i):
        x = seed % m or 1
        for _ in range(i + 1):          # discard(i) then operator()
            x = (a * x) % m
        return np.float32(np.float32(x - 1) / np.float32(2147483648.0))

    # the 10000th output of minstd_rand seeded with 1 is 399268537 (ISO C++ [rand.predef])
    x = 1
    for _ in range(10000):
        x = (a * x) % m
    assert x == 399268537
    assert orc.random_weight(1, 9999) == float(np.float32(np.float32(399268537 - 1) / np.float32(2147483648.0)))
    for seed in (0, 1, 7, 2147483647, 2147483648, 4294967295, 123456789):
        for i in (0, 1, 2, 31, 1000, 4097):
            assert orc.random_weight(seed, i) == float(ref(seed, i)), (seed, i)
    # large index: fast exponentiation path
    assert orc.random_weight(5, 10**11) == float(np.float32(np.float32((pow(a, 10**11 + 1, m) * 5) % m - 1) / np.float32(2147483648.0)))
    u = np.array([orc.random_weight(42, i) for i in range(20000)])
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01


def test_oracle_monotone_constraints_hold_and_zero_list_still_switches_the_gain_formula():
    """TreeEvaluator restated (split_evaluator.h:25-205): along a +1 (-1) feature every left leaf is <= (>=) every
    right leaf; children inherit bounds through mid = (w_left + w_right) / 2."""
    import oracle as orc
    from tests.helpers import make_cuts, make_regression, squared_error_gpair
    x, y = make_regression(6000, 6, seed=3)
    cp, cv = make_cuts(x, 32)
    gb = orc.bin_dense(x, cp, cv)
    gp = squared_error_gpair(np.zeros(6000, np.float32), y)
    p = orc.Param.make(max_depth=5)
    mono = [1, -1, 0, 1, 0, -1]
    o = orc.update_tree(gb, cp, cv, gp, p, monotone=mono)
    nodes = o["nodes"]

    def leaves(i):
        n = nodes[i]
        return [n.leaf_value] if n.is_leaf else leaves(n.left) + leaves(n.right)
    seen = 0
    for n in nodes:
        if not n.is_leaf and mono[n.findex] != 0:
            lo, hi = leaves(n.left), leaves(n.right)
            assert (max(lo) <= min(hi)) if mono[n.findex] > 0 else (min(lo) >= max(hi))
            seen += 1
    assert seen > 0
    plain = orc.update_tree(gb, cp, cv, gp, p)
    zeros = orc.update_tree(gb, cp, cv, gp, p, monotone=[0] * 6)   # has_constraint_: weight-based gains
    assert [n.findex for n in plain["nodes"]] == [n.findex for n in zeros["nodes"]]  # same tree here ...
    assert any(a.loss_chg != b.loss_chg for a, b in zip(plain["nodes"], zeros["nodes"]))  # ... other gain arithmetic


def test_oracle_feature_set_restricts_splits_and_row_sampling_zeroes_gradients():
    import oracle as orc
    from tests.helpers import make_cuts, make_regression, squared_error_gpair
    x, y = make_regression(5000, 10, seed=4)
    cp, cv = make_cuts(x, 16)
    gb = orc.bin_dense(x, cp, cv)
    gp = squared_error_gpair(np.zeros(5000, np.float32), y)
    p = orc.Param.make(max_depth=4)
    fs = [1, 4, 7]
    o = orc.update_tree(gb, cp, cv, gp, p, feature_set=fs)
    assert {n.findex for n in o["nodes"] if not n.is_leaf} <= set(fs)
    s = orc.update_tree(gb, cp, cv, gp, p, subsample=0.5, sample_seed=77, want_hist=True)
    dropped = np.array([orc.random_weight(77, i) > np.float32(0.5) for i in range(5000)])
    assert 0.4 < dropped.mean() < 0.6
    assert np.array_equal(s["position"] < 0, dropped)             # zero hessian -> bit-complemented leaf id
    # the root histogram holds only the kept rows: its hessian total = kept rows * quantised 1.0
    kept = int((~dropped).sum())
    assert int(s["hist"][0][cp[0]:cp[1], 1].sum()) == kept * int(np.float64(1.0) / s["scale"][3])


def test_mt19937_helper_matches_the_standard_known_answers():
    """tests/helpers.MT19937 stands in for the Booster's std::mt19937 when fixtures need the per-tree sampler seeds:
    ISO C++ [rand.predef] fixes the 10000th output of a default-constructed engine (seed 5489) to 4123659995."""
    from tests.helpers import MT19937
    m = MT19937(5489)
    v = [m.next() for _ in range(10000)]
    assert v[0] == 3499211612 and v[-1] == 4123659995
    assert MT19937(0).next() == 2357136044          # first output for seed 0 (the reference's default `seed`)


def test_categorical_evaluation_known_answers(golden_dir):
    """Partition-based and one-hot categorical split evaluation restated in the oracle,
    pinned by the reference's own literal test vectors before any CUDA is written for it."""
    spec = json.load(open(os.path.join(golden_dir, "eval_categorical_known_answers.json")))
    scale = _dummy_scale()
    for case in spec["cases"]:
        p = orc.Param.make(**case["param"])
        hist = _to_fixed(case["hist"], scale)
        parent = _to_fixed([case["parent"]], scale, as_float32=False)[0]
        c, is_cat, thresh, cats = orc.eval_node_cat(hist, parent, case["cut_ptrs"], case["cut_values"], p, scale,
                                                    case["is_cat"], case["max_cat_to_onehot"])
        e = case["expect"]
        name = case["name"]
        if "findex" in e:
            assert c.findex == e["findex"], name
        if "dir" in e:
            assert c.dir == e["dir"], name
        if "cats" in e:
            assert cats == set(e["cats"]), (name, cats)
        if "thresh" in e:
            assert thresh == e["thresh"], name
        if e.get("is_cat"):
            assert is_cat, name
        if e.get("fvalue_is_nan"):
            assert np.isnan(c.fvalue), name
        if e.get("sum_is_parent"):
            assert (c.left_sum.g + c.right_sum.g, c.left_sum.h + c.right_sum.h) == tuple(parent), name
        if "loss_chg" in e:
            assert abs(c.loss_chg - e["loss_chg"]) <= e.get("loss_chg_tol", 1e-6), (name, c.loss_chg)
        if "left_hess" in e:
            assert c.left_sum.h * scale[3] == e["left_hess"] and c.right_sum.h * scale[3] == e["right_hess"], name


def test_categorical_partition_finds_the_best_subset():
    """The property behind the reference's TestPartitionBasedSplit (tests/cpp/tree/test_evaluate_splits.h:25-45,
    test_evaluate_splits.cc SetUp: 'enumerate all possible partitions'): scanning the categories in the order of
    their leaf weights finds the best of ALL 2^n - 2 subsets.  Checked for the oracle on random 6-category histograms
    without missing values."""
    import itertools
    scale = _dummy_scale()
    p = orc.Param.make(min_child_weight=0.0, reg_lambda=0.0)
    rng = np.random.default_rng(7)
    for trial in range(40):
        g = rng.uniform(-4, 4, size=6).round(3)
        h = rng.uniform(0.5, 3, size=6).round(3)
        hist = _to_fixed(np.stack([g, h], axis=1).tolist(), scale)
        parent = hist.sum(axis=0)
        c, is_cat, thresh, cats = orc.eval_node_cat(hist, parent, [0, 6], [0.0, 1.0, 2.0, 3.0, 4.0, 5.0], p, scale, [1], 0)
        gq, hq = hist[:, 0].astype(np.float64) * scale[2], hist[:, 1].astype(np.float64) * scale[3]
        best = -np.inf
        for k in range(1, 6):
            for sub in itertools.combinations(range(6), k):
                m = np.zeros(6, dtype=bool)
                m[list(sub)] = True
                gain = gq[m].sum() ** 2 / hq[m].sum() + gq[~m].sum() ** 2 / hq[~m].sum()
                best = max(best, gain)
        parent_gain = gq.sum() ** 2 / hq.sum()
        assert is_cat and 0 < len(cats) < 6
        assert abs((c.loss_chg + np.float32(parent_gain)) - best) <= 1e-4 * max(1.0, abs(best)), (trial, c.loss_chg, best)
        # the reported category set holds the categories that go right, whichever scan direction won
        right = sorted(cats)
        assert c.right_sum.g == int(hist[right, 0].sum()) and c.right_sum.h == int(hist[right, 1].sum()), trial


# ---------------------------------------------------------------------------------------------------------------------
# Interaction constraints (src/tree/constraints.cu), SURVEY 8(f) rank 3: oracle first, the CUDA path comes next.
# ---------------------------------------------------------------------------------------------------------------------
def _ic_cases():
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "interaction_constraints_known_answers.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", _ic_cases(), ids=lambda c: c["name"])
def test_interaction_constraints_known_answers(case):
    """Split / QueryNode / Query of the oracle against the literal expectations of the reference's unit tests
    (tests/cpp/tree/test_constraints.cu:127-318)."""
    ic = orc.InteractionConstraints(case["sets"], case["n_features"], max_nodes=16)
    for step in case["steps"]:
        if "split" in step:
            ic.split(*step["split"])
        for nid, want in step.get("node_bits", {}).items():
            assert ic.query_node(int(nid)) == want, (case["name"], nid)
        for nid, want in step.get("query_node", {}).items():
            assert ic.query_node(int(nid)) == want, (case["name"], nid)
        if "query" in step:
            for nid, want in step["query"]["answers"].items():
                assert ic.query(step["query"]["list"], int(nid)) == want, (case["name"], nid)


def test_interaction_query_sorts_and_intersects():
    """Query answers in ascending feature order whatever the order of the sampled list (copy_if over a counting
    iterator, constraints.cu:222-231) and keeps only sampled features."""
    ic = orc.InteractionConstraints([[0, 3, 5], [1, 2]], 6, max_nodes=8)
    ic.split(0, 5, 1, 2)
    assert ic.query([5, 0, 4, 3], 1) == [0, 3, 5]
    assert ic.query([4, 5], 2) == [5]
    assert ic.query([4, 1], 1) == []
    assert ic.query([5, 0, 4, 3], 0) == [5, 0, 4, 3]          # the root is never filtered nor reordered


def _paths(nodes):
    out = []

    def walk(i, feats):
        if nodes[i].is_leaf:
            out.append(feats)
            return
        walk(nodes[i].left, feats + [nodes[i].findex])
        walk(nodes[i].right, feats + [nodes[i].findex])
    walk(0, [])
    return out


@pytest.mark.parametrize("shared", [False, True], ids=["per_query", "shared_result_buffer"])
def test_tree_respects_interaction_constraints(shared):
    """Whole-tree restatement: with per-query answers every root-to-leaf path uses features of one constraint set
    (or a single free feature); the unconstrained tree on the same data mixes them, so the constraint is what
    shapes the tree.  With the shared result buffer (what the reference's evaluation kernel reads) a tree is still
    produced and its first two levels agree with the per-query tree: the root is unfiltered and a level of one batch
    with equal answers aliases harmlessly."""
    rng = np.random.default_rng(5)
    n, f = 4000, 6
    x = rng.normal(size=(n, f)).astype(np.float32)
    y = (x[:, 0] * x[:, 3] + x[:, 1] * x[:, 4] + 0.5 * x[:, 2] * x[:, 5]).astype(np.float32)
    from tests.helpers import make_cuts
    cut_ptrs, cut_values = make_cuts(x, 16)
    gbins = orc.bin_dense(x, cut_ptrs, cut_values)
    gpair = np.stack([-y, np.ones_like(y)], axis=1).astype(np.float32)
    p = orc.Param.make(max_depth=4, eta=0.5, min_child_weight=1.0)
    sets = [[0, 1, 2], [3, 4, 5]]
    free = orc.update_tree(gbins, cut_ptrs, cut_values, gpair, p)
    mixed = [pa for pa in _paths(free["nodes"]) if any(a < 3 for a in pa) and any(a >= 3 for a in pa)]
    assert mixed, "test data must make the unconstrained tree cross the sets"
    t = orc.update_tree(gbins, cut_ptrs, cut_values, gpair, p, interaction=sets, interaction_shared_buffer=shared)
    assert len(t["nodes"]) > 3
    if not shared:
        for pa in _paths(t["nodes"]):
            assert all(a < 3 for a in pa) or all(a >= 3 for a in pa), pa
    else:
        ref = orc.update_tree(gbins, cut_ptrs, cut_values, gpair, p, interaction=sets)
        for i in range(3):
            assert t["nodes"][i].findex == ref["nodes"][i].findex and t["nodes"][i].fvalue == ref["nodes"][i].fvalue
    # the root is evaluated on every feature either way
    assert t["nodes"][0].findex == free["nodes"][0].findex and t["nodes"][0].fvalue == free["nodes"][0].fvalue


def test_oracle_node_feature_sets_follow_the_evaluation_order():
    """colsample_bynode restated: the k-th node evaluated (root, then left and right child of every valid candidate in
    Pop order) takes the k-th draw of the tree's stream; every split feature comes from its node's draw and the number
    of draws used is the number of evaluated nodes."""
    from tests.helpers import make_cuts, make_regression, squared_error_gpair
    x, y = make_regression(8000, 10, seed=5)
    cp, cv = make_cuts(x, 32)
    gb = orc.bin_dense(x, cp, cv)
    gp = squared_error_gpair(np.zeros_like(y), y)
    rng = np.random.default_rng(1)
    sets = np.stack([np.sort(rng.permutation(10)[:4]) for _ in range(31)]).astype(np.int32)
    for params in (dict(max_depth=4), dict(max_depth=0, max_leaves=8, grow_policy="lossguide")):
        o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), node_feature_sets=sets, max_nodes=31)
        nodes = o["nodes"]
        # replay the evaluation order from the finished tree
        if params.get("grow_policy") == "lossguide":   # children are evaluated when their parent is expanded
            order = [0]
            for nd in sorted((n for n in nodes if not n.is_leaf), key=lambda n: n.left):
                order += [nd.left, nd.right]
            order = order[:o["eval_count"]]
        else:
            order, cur = [0], [0]
            while cur:
                nxt = []
                for i in cur:
                    if not nodes[i].is_leaf and nodes[i].depth + 1 < params["max_depth"]:
                        nxt += [nodes[i].left, nodes[i].right]
                order += nxt
                cur = nxt
        assert o["eval_count"] == len(order)
        for k, i in enumerate(order):
            if not nodes[i].is_leaf:
                assert nodes[i].findex in sets[k], (params, k, i)
    # one full-width draw per node changes nothing
    full = np.tile(np.arange(10, dtype=np.int32), (31, 1))
    a = orc.update_tree(gb, cp, cv, gp, orc.Param.make(max_depth=4), node_feature_sets=full)
    b = orc.update_tree(gb, cp, cv, gp, orc.Param.make(max_depth=4))
    assert [(n.findex, n.fvalue) for n in a["nodes"]] == [(n.findex, n.fvalue) for n in b["nodes"]]


def test_gradient_based_sampling_properties():
    """GradientBasedSampling::Sample restated: the properties the reference's own (statistical) tests check,
    tests/cpp/tree/gpu_hist/test_sampler.cu:79-179 -- the bit-level pin is tests/test_reference_golden.py, on trees of the
    unmodified reference.  The threshold mu makes the EXPECTED number of sampled rows equal
    n * subsample (the CpuGpuConsistency property, :129-179), rows at or above mu are kept untouched, kept rows below
    it are scaled by 1 / p so that the gradient sum stays unbiased, dropped rows are zero."""
    from tests.helpers import make_regression, squared_error_gpair
    _, y = make_regression(20000, 4, seed=9)
    gp = squared_error_gpair(np.zeros_like(y), y)
    scale, _ = orc.quantiser(gp)
    gq = orc.quantise(gp, scale)
    for subsample in (0.1, 0.5, 0.9):
        out, mu, tidx, rag = orc.gradient_based_sample(gq, scale, subsample, seed=1234)
        p = np.minimum(1.0, rag / np.float32(mu))
        assert abs(float(p.sum()) - int(np.float32(len(y)) * np.float32(subsample))) <= 1.0, (subsample, p.sum())
        kept = (out[:, 1] != 0)
        assert abs(kept.sum() - p.sum()) <= 4 * np.sqrt((p * (1 - p)).sum()) + 1      # Poisson sampling spread
        sure = rag >= np.float32(mu)
        assert np.array_equal(out[sure], gq[sure])
        assert (out[~kept] == 0).all()
        scaled = kept & ~sure
        ratio = out[scaled, 1].astype(np.float64) / gq[scaled, 1].astype(np.float64)
        assert np.allclose(ratio, 1.0 / p[scaled], rtol=1e-5)
        # unbiased: the gradient sum is preserved up to sampling noise
        g_before, g_after = gq[:, 0].sum() * scale[2], out[:, 0].sum() * scale[2]
        noise = np.sqrt((((gq[:, 0] * scale[2]) ** 2) * (1 - p) / np.maximum(p, 1e-12)).sum())
        assert abs(g_after - g_before) <= 5 * noise + 1e-6
    # the same trial stream as uniform sampling: RandomWeight(seed)(i)
    out1, *_ = orc.gradient_based_sample(gq, scale, 0.5, seed=7)
    out2, *_ = orc.gradient_based_sample(gq, scale, 0.5, seed=7)
    assert np.array_equal(out1, out2)
    # and a whole tree grows on the sampled pairs
    from tests.helpers import make_cuts
    x, y2 = make_regression(6000, 5, seed=10)
    cp, cv = make_cuts(x, 16)
    gb = orc.bin_dense(x, cp, cv)
    gp2 = squared_error_gpair(np.zeros_like(y2), y2)
    t = orc.update_tree(gb, cp, cv, gp2, orc.Param.make(max_depth=4), subsample=0.3, sample_seed=5,
                        sampling_method="gradient_based")
    u = orc.update_tree(gb, cp, cv, gp2, orc.Param.make(max_depth=4), subsample=0.3, sample_seed=5)
    assert len(t["nodes"]) >= 3 and (t["position"] < 0).sum() > 0
    assert (t["position"] < 0).sum() != (u["position"] < 0).sum()


def test_interaction_constraint_oracle_against_a_set_model():
    """orc_ic_split / orc_ic_query (bit fields, CSR sets) against the three-line set model of the reference's kernels
    on random constraint sets, split sequences and sampled lists, incl. features beyond one 64-bit word."""
    rng = np.random.default_rng(123)
    for trial in range(60):
        nf = int(rng.integers(3, 200))
        sets = [sorted(set(int(v) for v in rng.integers(0, nf, size=rng.integers(1, 6)))) for _ in range(rng.integers(1, 7))]
        ic = orc.InteractionConstraints(sets, nf, max_nodes=64)
        model = [set() for _ in range(64)]
        n_nodes = 1
        for _ in range(int(rng.integers(1, 20))):
            if n_nodes + 2 > 64:
                break
            node = int(rng.integers(0, n_nodes))
            f = int(rng.integers(0, nf))
            left, right = n_nodes, n_nodes + 1
            n_nodes += 2
            ic.split(node, f, left, right)
            for s in sets:
                if f in s:
                    model[node] |= set(s)
            model[left] |= model[node]
            model[right] |= model[node]
            for k in (node, left, right):
                model[k].add(f)
            for k in range(n_nodes):
                assert ic.query_node(k) == sorted(model[k]), (trial, k)
            sampled = [int(v) for v in rng.permutation(nf)[:int(rng.integers(1, nf + 1))]]
            k = int(rng.integers(0, n_nodes))
            want = sampled if k == 0 else sorted(model[k] & set(sampled))
            assert ic.query(sampled, k) == want, (trial, k)


def test_thrust_shuffle_known_answers(golden_dir):
    """orc_thrust_shuffle against permutations produced by the toolkit's own thrust::shuffle (host backend of the same
    generic algorithm the device runs; tools/thrust_vectors/shuffle_vectors.cpp): sizes around the cipher's
    power-of-two boundaries (<= 16 uses 4 bits), seeds 0 / 2^31-1 / 2^32-1 (the engine's seed folding)."""
    spec = json.load(open(os.path.join(golden_dir, "thrust_shuffle_known_answers.json")))
    assert len(spec["cases"]) >= 50
    for c in spec["cases"]:
        assert orc.thrust_shuffle(c["seed"], c["n"]).tolist() == c["perm"], (c["seed"], c["n"])
    # ColSample: a sorted subset of the right size, the whole set at 1.0, never empty
    feats = np.arange(40, dtype=np.int32) * 2
    s = orc.col_sample(feats, 0.5, 12345)
    perm = orc.thrust_shuffle(12345, 40)
    assert s.tolist() == sorted(feats[perm[:20]].tolist())
    assert orc.col_sample(feats, 1.0, 1).tolist() == feats.tolist()
    assert orc.col_sample(feats, 0.001, 1).size == 1
