"""Kernel SOURCES that have not run on a GPU yet, executed on the CPU.

tests/emu/cuda_emu.h is a small CUDA emulator (one OS thread per CUDA thread, barriers for __syncthreads and the warp
collectives, statics for shared memory).  The harnesses under tests/emu/ #include the real .cu files of
xgboost_b200/csrc with B200_KERNELS_ONLY (host launchers compiled out) and run their kernels on inputs written here;
the outputs are compared with the oracle.  This checks the kernel text -- indexing, control flow, arithmetic, data flow
through shared memory and shuffles -- not the hardware (memory model, scheduling, device-only arithmetic).
"""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle as orc
from tests.conftest import ROOT
from tests.helpers import make_cuts, make_regression, squared_error_gpair

EMU = os.path.join(ROOT, "tests", "emu")
BUILD = os.path.join(ROOT, "build", "emu")


# B200_EMU_SANITIZE=address|thread|undefined builds the harnesses with that sanitizer into their own directory: the
# kernels' global buffers are heap vectors and their __shared__ arrays function statics, so an out-of-bounds index in
# kernel text is an ASan report (and a non-zero exit) instead of a silently wrong neighbour.
SANITIZE = os.environ.get("B200_EMU_SANITIZE", "")


def _build(name):
    build = BUILD + ("_" + SANITIZE if SANITIZE else "")
    os.makedirs(build, exist_ok=True)
    exe = os.path.join(build, name)
    srcs = [os.path.join(EMU, name + ".cc"), os.path.join(EMU, "cuda_emu.cc")]
    csrc = os.path.join(ROOT, "xgboost_b200", "csrc")
    deps = srcs + [os.path.join(EMU, "cuda_emu.h")] + \
        [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        # -fno-sanitize-recover: a UBSan finding ends the harness non-zero instead of printing and carrying on
        extra = ["-fsanitize=" + SANITIZE, "-fno-sanitize-recover=all", "-g", "-fno-omit-frame-pointer"] if SANITIZE else []
        subprocess.run(["g++", "-std=c++17", "-O1", "-pthread"] + extra +
                       ["-I" + os.path.join(EMU, "include"), "-I" + EMU,
                        "-I" + os.path.join(ROOT, "include"), "-I" + csrc, "-o", exe] + srcs, check=True)
    return exe


def _run(exe, mode, workdir):
    subprocess.run([exe, mode, workdir], check=True, timeout=900)


@pytest.mark.parametrize("n,colsample", [(1, 0.5), (7, 0.5), (16, 0.3), (17, 0.9), (256, 0.5), (257, 0.25), (1000, 0.02)])
def test_col_sample_stream_kernel_source(tmp_path, n, colsample):
    """ColSampleStreamKernel (csrc/colsample.cu) = ColSample restated in the oracle, for every draw of a stream."""
    exe = _build("emu_colsample_predict")
    rng = np.random.default_rng(n)
    features = np.sort(rng.choice(3 * n + 5, size=n, replace=False)).astype(np.int32)
    seeds = np.array([0, 1, 2147483647, 4294967295] + [int(v) for v in rng.integers(0, 2 ** 32, size=8)], dtype=np.uint32)
    n_sel = max(1, int(np.float32(colsample) * np.float32(n)))
    d = str(tmp_path)
    features.tofile(os.path.join(d, "features.bin"))
    seeds.tofile(os.path.join(d, "seeds.bin"))
    np.array([n_sel], dtype=np.int32).tofile(os.path.join(d, "meta.bin"))
    _run(exe, "colsample", d)
    out = np.fromfile(os.path.join(d, "out.bin"), dtype=np.int32).reshape(len(seeds), n_sel)
    for k, seed in enumerate(seeds):
        assert np.array_equal(out[k], orc.col_sample(features, colsample, int(seed))), (n, colsample, int(seed))


PRED_NODE = np.dtype([("findex", "<i4"), ("split_lbin", "<i4"), ("left", "<i4"), ("right", "<i4"),
                      ("leaf_value", "<f4"), ("default_left", "<i4")])


def _forest_arrays(trees, cp, cv, bin_words):
    """What b200hist_forest_add_tree builds on the host: nodes with bin thresholds, roots, local-bin sets."""
    nodes, roots, bits = [], [], []
    for t in trees:
        base = len(nodes)
        roots.append(base)
        is_cat = t.get("is_cat")
        for i, nd in enumerate(t["nodes"]):
            w = np.zeros(bin_words, dtype=np.uint32)
            if nd.is_leaf:
                nodes.append((-1, 0, base + i, base + i, nd.leaf_value, int(nd.default_left)))
            else:
                beg, end = int(cp[nd.findex]), int(cp[nd.findex + 1])
                if is_cat is not None and is_cat[i]:
                    lbin = -1
                    for b in range(beg, end):
                        if int(cv[b]) in t["cat_sets"][i]:
                            w[(b - beg) >> 5] |= np.uint32(1 << ((b - beg) & 31))
                else:
                    # 1 + number of the feature's cuts < the condition (csrc/predict.cu add_tree)
                    lbin = (1 + int(np.searchsorted(cv[beg:end], np.float32(nd.fvalue), side="left"))
                            if np.float32(nd.fvalue) > -np.inf else 0)
                nodes.append((nd.findex, lbin, base + nd.left, base + nd.right, nd.leaf_value, int(nd.default_left)))
            bits.append(w)
    return np.array(nodes, dtype=PRED_NODE), np.array(roots, dtype=np.uint32), np.stack(bits)


def _page_bytes(gb, cp, bin_bytes, null_symbol):
    """Our dense page layout: feature-local symbols, rows padded to whole 32-feature blocks."""
    n, f = gb.shape
    width = (f + 31) // 32 * 32
    sym = np.where(gb >= 0, gb - cp[:-1].astype(np.int64)[None, :], null_symbol)
    page = np.zeros((n, width), dtype=np.uint8 if bin_bytes == 1 else np.uint16)
    page[:, :f] = sym
    return page.view(np.uint8).reshape(n, width * bin_bytes)


@pytest.mark.parametrize("kind", ["numerical_u8", "missing_u16", "categorical", "sparse", "foreign_cuts"])
def test_predict_kernel_sources(tmp_path, kind):
    """PredictMarginKernel / PredictLeafKernel (csrc/predict.cu) against the oracle's walk and tree-order sum."""
    exe = _build("emu_colsample_predict")
    if kind == "categorical":
        from tests.helpers import make_categorical_data
        from tests.test_reference_golden import categorical_global_bins
        case = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_cpu_categorical_trees.json")))["cases"][0]
        case = dict(case, n=3000)
        x, y, types = make_categorical_data(case)
        is_cat = [1 if t == "c" else 0 for t in types]
        cp = np.array(case["cut_ptrs"], dtype=np.uint32)
        cv = np.array(case["cut_values_bits"], dtype=np.uint32).view(np.float32)
        gb = categorical_global_bins(x, cp, cv, is_cat)
        kw = dict(feature_is_cat=is_cat)
    else:
        x, y = make_regression(3000, 9, seed=5, missing={"missing_u16": 0.1, "sparse": 0.6}.get(kind, 0.0))
        cp, cv = make_cuts(x, 300 if kind == "missing_u16" else 32)
        gb = orc.bin_dense(x, cp, cv)
        kw = {}
    pred = np.zeros(len(y), dtype=np.float32)
    trees = []
    for _ in range(4):
        o = orc.update_tree(gb, cp, cv, squared_error_gpair(pred, y), orc.Param.make(max_depth=4, eta=0.5), **kw)
        trees.append(o)
        pred = orc.update_prediction_cache(o["nodes"], o["position"], pred)
    if kind == "foreign_cuts":
        # the forest meets a page quantised with other cuts: its conditions are no longer cuts of the page
        x, y = make_regression(3000, 9, seed=6, missing=0.05)
        cp, cv = make_cuts(x, 20)
        gb = orc.bin_dense(x, cp, cv)
        for o in trees:
            o["nodes"][0].fvalue = -np.inf if o is trees[0] else o["nodes"][0].fvalue   # nothing is below -inf
    max_bins = int(np.diff(cp).max())
    has_missing = bool((gb < 0).any())
    bin_bytes = 2 if max_bins + (1 if has_missing else 0) > 256 else 1
    null_symbol = max_bins if has_missing else 0xFFFFFFFF
    bin_words = (max_bins + 31) // 32
    nodes, roots, bits = _forest_arrays(trees, cp, cv, bin_words)
    page = _page_bytes(gb, cp, bin_bytes, null_symbol if has_missing else 0)
    row_stride = 0
    if kind == "sparse":
        row_stride = max(1, int((gb >= 0).sum(axis=1).max()))
        ell = np.full((len(y), row_stride), cv.size, dtype=np.uint32)
        for r in range(len(y)):
            present = gb[r][gb[r] >= 0]
            ell[r, :present.size] = present
        page = ell.view(np.uint8).reshape(len(y), row_stride * 4)
        bin_bytes, null_symbol = 4, cv.size
    base = np.full(len(y), 0.25, dtype=np.float32)
    d = str(tmp_path)
    np.array([len(y), gb.shape[1], page.shape[1], bin_bytes, null_symbol, len(trees), bin_words, row_stride],
             dtype=np.uint32).tofile(os.path.join(d, "meta.bin"))
    page.tofile(os.path.join(d, "bins.bin"))
    nodes.tofile(os.path.join(d, "nodes.bin"))
    roots.tofile(os.path.join(d, "roots.bin"))
    bits.tofile(os.path.join(d, "bin_bits.bin"))
    cp.tofile(os.path.join(d, "cut_ptrs.bin"))
    cv.tofile(os.path.join(d, "cut_values.bin"))
    base.tofile(os.path.join(d, "base.bin"))
    n_groups = 3
    tree_groups = np.array([(2 * t + 1) % n_groups for t in range(len(trees))], dtype=np.uint32)   # 1, 0, 2, 1
    base_groups = np.random.default_rng(3).normal(size=(len(y), n_groups)).astype(np.float32)
    tree_groups.tofile(os.path.join(d, "tree_groups.bin"))
    base_groups.tofile(os.path.join(d, "base_groups.bin"))
    from tests.helpers import pack_reference_page
    ref_buf, ref_stride, ref_symbols, ref_is_dense = pack_reference_page(gb, cp, dense_layout=kind != "sparse")
    ref_buf.tofile(os.path.join(d, "ref_page.bin"))
    np.array([ref_stride, ref_symbols, 0 if kind == "sparse" else 1, 1 if ref_is_dense else 0],
             dtype=np.uint32).tofile(os.path.join(d, "ref_meta.bin"))
    _run(exe, "predict", d)
    margin = np.fromfile(os.path.join(d, "margin.bin"), dtype=np.float32)
    leaves = np.fromfile(os.path.join(d, "leaves.bin"), dtype=np.int32).reshape(len(y), len(trees))
    want = orc.predict_margin(gb, cp, cv, trees, base=base)
    assert np.array_equal(margin.view(np.uint32), want.view(np.uint32))
    for t, o in enumerate(trees):
        assert np.array_equal(leaves[:, t], orc.predict_leaf(gb, cp, cv, o["nodes"], o.get("is_cat"), o.get("cat_sets")))
    # several output groups: leaves go straight onto the output, tree after tree; leaf ids as floats
    got_groups = np.fromfile(os.path.join(d, "margin_groups.bin"), dtype=np.float32).reshape(len(y), n_groups)
    want_groups = orc.predict_margin_groups(gb, cp, cv, trees, tree_groups, n_groups, base=base_groups)
    assert np.array_equal(got_groups.view(np.uint32), want_groups.view(np.uint32))
    fleaves = np.fromfile(os.path.join(d, "leaves_f32.bin"), dtype=np.float32).reshape(len(y), len(trees))
    assert np.array_equal(fleaves, leaves.astype(np.float32))
    # ... and straight from the page as the reference stores it (bit-packed, 6..10-bit symbols here)
    ref_margin = np.fromfile(os.path.join(d, "ref_margin.bin"), dtype=np.float32)
    ref_groups = np.fromfile(os.path.join(d, "ref_margin_groups.bin"), dtype=np.float32).reshape(len(y), n_groups)
    ref_leaves = np.fromfile(os.path.join(d, "ref_leaves.bin"), dtype=np.float32).reshape(len(y), len(trees))
    assert np.array_equal(ref_margin.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(ref_groups.view(np.uint32), want_groups.view(np.uint32))
    assert np.array_equal(ref_leaves, leaves.astype(np.float32))


# ---------------------------------------------------------------------------------------------------------------------
# Split evaluation (csrc/evaluate.cu) under the emulator.  The emulator divides in IEEE arithmetic like the oracle (the
# device uses div.approx), so everything -- loss_chg included -- must agree bit for bit.
# ---------------------------------------------------------------------------------------------------------------------
CAND = np.dtype([("loss_chg", "<f4"), ("dir", "<i4"), ("findex", "<i4"), ("fvalue", "<f4"), ("left_g", "<i8"),
                 ("left_h", "<i8"), ("right_g", "<i8"), ("right_h", "<i8"), ("base_weight", "<f4"),
                 ("left_weight", "<f4"), ("right_weight", "<f4"), ("split_gbin", "<i4")])


def _evaluate(tmp_path, cp, cv, hists, parents, params, scale, depth=None, feature_set=None, n_levels=0, ic_list=None,
              ic_count=None, ic_stride=0, mono=None, lower=None, upper=None, is_cat=None, onehot=4, threshold=64):
    exe = _build("emu_evaluate")
    n_eval, total_bins = hists.shape[0], hists.shape[1]
    nf = cp.size - 1
    n_fs = nf if feature_set is None else (feature_set.shape[-1])
    cat_words = int(cv.max()) // 32 + 1
    max_cat_bins = int(max([cp[f + 1] - cp[f] for f in range(nf) if is_cat is not None and is_cat[f]] or [1]))
    ic_mode = 0 if ic_count is None else (2 if ic_stride > 0 else 1)
    d = str(tmp_path)
    np.array([nf, total_bins, n_eval, n_fs, 0 if feature_set is None else 1, n_levels, ic_mode, ic_stride,
              0 if mono is None else 1, 0 if is_cat is None else 1, onehot, threshold, cat_words,
              (max_cat_bins + 31) // 32, max_cat_bins], dtype=np.int32).tofile(os.path.join(d, "meta.bin"))
    p = orc.Param.make(**params)
    np.array([p.learning_rate, p.reg_lambda, p.reg_alpha, p.min_child_weight, p.max_delta_step, p.min_split_loss],
             dtype=np.float32).tofile(os.path.join(d, "param_f.bin"))
    np.array([p.max_depth, p.max_leaves, p.grow_policy], dtype=np.int32).tofile(os.path.join(d, "param_i.bin"))
    np.asarray(scale, dtype=np.float64).tofile(os.path.join(d, "quant.bin"))
    cp.astype(np.uint32).tofile(os.path.join(d, "cut_ptrs.bin"))
    cv.astype(np.float32).tofile(os.path.join(d, "cut_values.bin"))
    np.ascontiguousarray(hists, dtype=np.int64).tofile(os.path.join(d, "hist.bin"))
    np.asarray(parents, dtype=np.int64).tofile(os.path.join(d, "parents.bin"))
    np.asarray(depth if depth is not None else np.zeros(n_eval), dtype=np.int32).tofile(os.path.join(d, "depth.bin"))
    for name, arr, dt in (("feature_set", feature_set, np.int32), ("ic_list", ic_list, np.int32),
                          ("ic_count", ic_count, np.int32), ("mono", mono, np.int32), ("is_cat", is_cat, np.uint8)):
        if arr is not None:
            np.ascontiguousarray(arr, dtype=dt).tofile(os.path.join(d, name + ".bin"))
    np.asarray(lower if lower is not None else np.full(n_eval, -3.402823466e+38), dtype=np.float32).tofile(os.path.join(d, "lower.bin"))
    np.asarray(upper if upper is not None else np.full(n_eval, 3.402823466e+38), dtype=np.float32).tofile(os.path.join(d, "upper.bin"))
    subprocess.run([exe, d], check=True, timeout=900)
    cand = np.fromfile(os.path.join(d, "cand.bin"), dtype=CAND)
    flags = np.fromfile(os.path.join(d, "node_is_cat.bin"), dtype=np.uint8)
    bits = np.fromfile(os.path.join(d, "cat_bits.bin"), dtype=np.uint32).reshape(n_eval, -1)
    return cand, flags, bits


def _same_candidate(c, o, name, cat=False):
    assert (int(c["findex"]), int(c["dir"])) == (o.findex, o.dir), name
    assert (int(c["left_g"]), int(c["left_h"]), int(c["right_g"]), int(c["right_h"])) == \
        (o.left_sum.g, o.left_sum.h, o.right_sum.g, o.right_sum.h), name
    for a, b in ((c["loss_chg"], o.loss_chg), (c["base_weight"], o.base_weight), (c["left_weight"], o.left_weight),
                 (c["right_weight"], o.right_weight)):
        assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32), (name, a, b)
    if cat:
        assert np.isnan(c["fvalue"]) and c["split_gbin"] == -1, name
    else:
        assert np.float32(c["fvalue"]).view(np.uint32) == np.float32(o.fvalue).view(np.uint32), name
        assert int(c["split_gbin"]) == o.split_gbin, name


def _histograms(n=6000, f=14, max_bin=48, missing=0.1, seed=3):
    from tests.helpers import logistic_gpair
    x, y = make_regression(n, f, seed=seed, missing=missing)
    cp, cv = make_cuts(x, max_bin)
    gb = orc.bin_dense(x, cp, cv)
    gp = logistic_gpair(y, (y > 0).astype(np.float32))
    scale, _ = orc.quantiser(gp)
    gq = orc.quantise(gp, scale)
    rng = np.random.default_rng(seed)
    subsets = [np.arange(n), rng.choice(n, 900, replace=False), rng.choice(n, 70, replace=False),
               rng.choice(n, 2500, replace=False), rng.choice(n, 9, replace=False)]
    hists = np.stack([orc.build_hist(gb, gq, s.astype(np.uint32), cv.size) for s in subsets])
    parents = np.array([gq[s].sum(axis=0) for s in subsets], dtype=np.int64)
    return cp, cv, hists, parents, scale


@pytest.mark.parametrize("params", [dict(), dict(reg_alpha=0.4, min_child_weight=20.0), dict(max_delta_step=0.6, reg_lambda=0.0)],
                         ids=["default", "alpha_mcw", "delta_step"])
def test_evaluation_kernel_sources_plain_and_feature_lists(tmp_path, params):
    cp, cv, hists, parents, scale = _histograms()
    nf = cp.size - 1
    p = orc.Param.make(**params)
    # all features
    cand, _, _ = _evaluate(tmp_path, cp, cv, hists, parents, params, scale)
    for k in range(len(parents)):
        _same_candidate(cand[k], orc.eval_node(hists[k], parents[k], cp, cv, p, scale), "all features, node %d" % k)
    # one list per evaluation slot (colsample_bynode): IcView with a stride
    rng = np.random.default_rng(8)
    lists = np.stack([np.sort(rng.permutation(nf)[:5]) for _ in range(len(parents))]).astype(np.int32)
    counts = np.full(len(parents), 5, dtype=np.int32)
    cand, _, _ = _evaluate(tmp_path, cp, cv, hists, parents, params, scale, ic_list=lists, ic_count=counts, ic_stride=5)
    for k in range(len(parents)):
        _same_candidate(cand[k], orc.eval_node(hists[k], parents[k], cp, cv, p, scale, feature_set=lists[k]),
                        "per-slot list, node %d" % k)
    # one shared list, a count per slot (interaction constraints): duplicates and an empty answer included
    shared = np.array([2, 3, 7, 7, 11, 12, 13, 0, 0, 0, 0, 0, 0, 0], dtype=np.int32)
    counts = np.array([7, 4, 0, 5, 1], dtype=np.int32)
    cand, _, _ = _evaluate(tmp_path, cp, cv, hists, parents, params, scale, ic_list=shared, ic_count=counts, ic_stride=0)
    for k in range(len(parents)):
        if counts[k] == 0:
            assert cand[k]["findex"] == -1 and cand[k]["loss_chg"] < -1e30, k
            continue
        _same_candidate(cand[k], orc.eval_node(hists[k], parents[k], cp, cv, p, scale, feature_set=shared[:counts[k]]),
                        "shared list, node %d" % k)
    # colsample_bylevel: one list per depth, the last row serves deeper nodes
    levels = np.stack([np.sort(rng.permutation(nf)[:6]) for _ in range(3)]).astype(np.int32)
    depth = np.array([0, 1, 2, 5, 1], dtype=np.int32)
    cand, _, _ = _evaluate(tmp_path, cp, cv, hists, parents, params, scale, depth=depth, feature_set=levels, n_levels=3)
    for k in range(len(parents)):
        _same_candidate(cand[k], orc.eval_node(hists[k], parents[k], cp, cv, p, scale, feature_set=levels[min(depth[k], 2)]),
                        "level list, node %d" % k)


def test_evaluation_kernel_sources_monotone(tmp_path):
    cp, cv, hists, parents, scale = _histograms(seed=4)
    nf = cp.size - 1
    mono = np.array([1, -1, 0, 1, 0, -1, 1, 0, 0, -1, 1, 1, 0, -1], dtype=np.int32)[:nf]
    lower = np.array([-3.402823466e+38, -0.2, -3.402823466e+38, 0.05, -1.0], dtype=np.float32)
    upper = np.array([3.402823466e+38, 3.402823466e+38, 0.1, 0.4, 1.0], dtype=np.float32)
    params = dict(reg_lambda=1.0)
    cand, _, _ = _evaluate(tmp_path, cp, cv, hists, parents, params, scale, mono=mono, lower=lower, upper=upper)
    p = orc.Param.make(**params)
    for k in range(len(parents)):
        o = orc.eval_node_mono(hists[k], parents[k], cp, cv, p, scale, monotone=mono, lower=float(lower[k]), upper=float(upper[k]))
        _same_candidate(cand[k], o, "monotone, node %d" % k)


@pytest.mark.parametrize("onehot,threshold", [(4, 64), (1, 64), (100, 64), (4, 5)])
def test_evaluation_kernel_sources_categorical(tmp_path, onehot, threshold):
    """CatFeatureScanKernel / CatFinishKernel incl. features with more than 32 and more than 64 categories (several
    arg-max chunks in both scan directions), next to numerical features, on nodes of very different sizes."""
    rng = np.random.default_rng(17)
    n = 6000
    n_cats = [2, 3, 5, 9, 31, 33, 70]
    cols, cut_ptrs, cut_values, is_cat = [], [0], [], []
    for k in n_cats:
        col = rng.integers(0, k, size=n).astype(np.float32)
        col[rng.random(n) < 0.08] = np.nan
        cols.append(col)
        cut_values += [float(v) for v in range(k)]
        cut_ptrs.append(len(cut_values))
        is_cat.append(1)
    xnum, _ = make_regression(n, 3, seed=23, missing=0.05)
    ncp, ncv = make_cuts(xnum, 32)
    for j in range(3):
        cols.append(xnum[:, j])
        cut_values += [float(v) for v in ncv[ncp[j]:ncp[j + 1]]]
        cut_ptrs.append(len(cut_values))
        is_cat.append(0)
    cp, cv = np.array(cut_ptrs, dtype=np.uint32), np.array(cut_values, dtype=np.float32)
    from tests.test_reference_golden import categorical_global_bins
    gb = categorical_global_bins(np.column_stack(cols).astype(np.float32), cp, cv, is_cat)
    effect = rng.normal(size=70)
    y = (effect[np.nan_to_num(cols[6], nan=0).astype(int)] + 0.5 * effect[np.nan_to_num(cols[4], nan=0).astype(int)] +
         np.nan_to_num(xnum[:, 0]) + 0.3 * rng.normal(size=n)).astype(np.float32)
    gp = squared_error_gpair(np.zeros(n, np.float32), y)
    scale, _ = orc.quantiser(gp)
    gq = orc.quantise(gp, scale)
    subsets = [np.arange(n), rng.choice(n, 900, replace=False), rng.choice(n, 60, replace=False)]
    hists = np.stack([orc.build_hist(gb, gq, s.astype(np.uint32), cv.size) for s in subsets])
    parents = np.array([gq[s].sum(axis=0) for s in subsets], dtype=np.int64)
    params = dict(min_child_weight=2.0)
    cand, flags, bits = _evaluate(tmp_path, cp, cv, hists, parents, params, scale, is_cat=np.array(is_cat, dtype=np.uint8),
                                  onehot=onehot, threshold=threshold)
    p = orc.Param.make(**params)
    n_cat_winners = 0
    for k in range(len(parents)):
        o, o_is_cat, _, o_cats = orc.eval_node_cat(hists[k], parents[k], cp, cv, p, scale, is_cat, onehot, threshold)
        _same_candidate(cand[k], o, "categorical, node %d" % k, cat=o_is_cat)
        assert bool(flags[k]) == o_is_cat, k
        if o_is_cat:
            n_cat_winners += 1
            got = {32 * w + (31 - b) for w in range(bits.shape[1]) for b in range(32) if (int(bits[k, w]) >> b) & 1}
            assert got == o_cats, (k, got, o_cats)
    if threshold == 64 and onehot != 100:   # in the other configurations a numerical feature wins on this data
        assert n_cat_winners >= 1


# ---------------------------------------------------------------------------------------------------------------------
# Control kernels of csrc/updater.cu and the categorical re-binning kernel of csrc/ellpack.cu
# ---------------------------------------------------------------------------------------------------------------------
def test_node_set_level_kernel_source(tmp_path):
    """NodeSetLevelKernel deals draw seq + j to evaluation slot j and moves the stream position, step after step."""
    exe = _build("emu_updater")
    rng = np.random.default_rng(3)
    n_sets, n_per = 40, 7
    sets = rng.integers(0, 100, size=(n_sets, n_per)).astype(np.int32)
    steps = np.array([1, 2, 4, 0, 6, 2], dtype=np.int32)
    d = str(tmp_path)
    np.array([n_sets, n_per, 0], dtype=np.int32).tofile(os.path.join(d, "meta.bin"))
    sets.tofile(os.path.join(d, "sets.bin"))
    steps.tofile(os.path.join(d, "steps.bin"))
    _run(exe, "nodeset", d)
    lists = np.fromfile(os.path.join(d, "lists.bin"), dtype=np.int32).reshape(-1, n_per)
    counts = np.fromfile(os.path.join(d, "counts.bin"), dtype=np.int32)
    seq = int(np.fromfile(os.path.join(d, "seq.bin"), dtype=np.int32)[0])
    assert seq == steps.sum() and lists.shape[0] == steps.sum()
    assert np.array_equal(lists, sets[:steps.sum()]) and (counts == n_per).all()


@pytest.mark.parametrize("with_level_set", [False, True], ids=["all_features", "level_set"])
def test_interaction_level_kernel_source(tmp_path, with_level_set):
    """IcLevelKernel: Split of the level's applied splits, per-slot counts and the shared Query buffer image, against
    the oracle's bit fields and a literal replay of the reference's sequential buffer writes.  130 features = three
    64-bit words; several levels are chained so that bit fields inherited from earlier launches are used."""
    exe = _build("emu_updater")
    rng = np.random.default_rng(11)
    nf = 130
    sets = [[0, 1, 129], [129, 64, 65], [65, 100], [2, 3], [3, 64], [100, 7, 8, 9]]
    set_ptr = np.cumsum([0] + [len(s) for s in sets]).astype(np.int32)
    set_feat = np.array([f for s in sets for f in s], dtype=np.int32)
    max_nodes = 63
    words = (nf + 63) // 64
    bits = np.zeros((max_nodes, words), dtype=np.uint64)
    ic = orc.InteractionConstraints(sets, nf, max_nodes=max_nodes)
    level_sets = np.stack([np.sort(rng.permutation(nf)[:60]) for _ in range(4)]).astype(np.int32)
    depth = np.zeros(max_nodes, dtype=np.int32)
    level, n_nodes = [0], 1
    pool = [0, 1, 129, 64, 65, 100, 2, 3, 7, 50, 77]      # split features: members of sets and free ones
    for d_level in range(4):
        parts = []
        for nidx in level:
            left, right = n_nodes, n_nodes + 1
            n_nodes += 2
            depth[left] = depth[right] = d_level + 1
            parts.append((nidx, int(rng.choice(pool)), left, right))
        dd = str(tmp_path)
        n_fs = level_sets.shape[1] if with_level_set else 0
        np.array([nf, len(sets), max_nodes, len(parts), n_fs, 4 if with_level_set else 0, 1 if with_level_set else 0],
                 dtype=np.int32).tofile(os.path.join(dd, "meta.bin"))
        set_ptr.tofile(os.path.join(dd, "set_ptr.bin"))
        set_feat.tofile(os.path.join(dd, "set_feat.bin"))
        bits.tofile(os.path.join(dd, "bits.bin"))
        np.array(parts, dtype=np.int32).tofile(os.path.join(dd, "part.bin"))
        depth.tofile(os.path.join(dd, "depth.bin"))
        if with_level_set:
            level_sets.tofile(os.path.join(dd, "feature_set.bin"))
        _run(exe, "iclevel", dd)
        bits = np.fromfile(os.path.join(dd, "bits_out.bin"), dtype=np.uint64).reshape(max_nodes, words)
        got_list = np.fromfile(os.path.join(dd, "list.bin"), dtype=np.int32)
        got_count = np.fromfile(os.path.join(dd, "count.bin"), dtype=np.int32)
        # the oracle's fields after the same splits
        for nidx, f, left, right in parts:
            ic.split(nidx, f, left, right)
        for k in range(n_nodes):
            mine = [f for f in range(nf) if (int(bits[k, f >> 6]) >> (f & 63)) & 1]
            assert mine == ic.query_node(k), (d_level, k)
        # the reference's buffer: every Query writes its answer to the front, in the order left0, right0, left1, ...
        sampled = level_sets[min(d_level + 1, 3)].tolist() if with_level_set else list(range(nf))
        buffer = [-7] * nf
        order = [c for _, _, left, right in parts for c in (left, right)]
        for j, node in enumerate(order):
            answer = ic.query(sampled, node)
            buffer[:len(answer)] = answer
            assert got_count[j] == len(answer), (d_level, j)
        longest = max(got_count[:len(order)])
        assert got_list[:longest].tolist() == buffer[:longest], d_level
        level = order if d_level < 2 else order[:6]


@pytest.mark.parametrize("threshold", [64, 1000])
def test_evaluation_kernel_sources_with_a_thousand_categories(tmp_path, threshold):
    """The categorical kernels at the width the driver allows (kMaxCatBins = 1024): a 1000-category feature is 32
    arg-max chunks per scan direction and a warp sort over 1000 weights."""
    rng = np.random.default_rng(41)
    n, k = 30000, 1000
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
    subsets = [np.arange(n), rng.choice(n, 4000, replace=False)]
    hists = np.stack([orc.build_hist(gb, gq, s.astype(np.uint32), cv.size) for s in subsets])
    parents = np.array([gq[s].sum(axis=0) for s in subsets], dtype=np.int64)
    params = dict(max_depth=6)
    cand, flags, bits = _evaluate(tmp_path, cp, cv, hists, parents, params, scale, is_cat=np.array(is_cat, dtype=np.uint8),
                                  onehot=4, threshold=threshold)
    p = orc.Param.make(**params)
    for j in range(len(parents)):
        o, o_is_cat, _, o_cats = orc.eval_node_cat(hists[j], parents[j], cp, cv, p, scale, is_cat, 4, threshold)
        assert o_is_cat and bool(flags[j]), j
        _same_candidate(cand[j], o, "1000 categories, node %d" % j, cat=True)
        got = {32 * w + (31 - b) for w in range(bits.shape[1]) for b in range(32) if (int(bits[j, w]) >> b) & 1}
        assert got == o_cats, (j, len(got), len(o_cats))


@pytest.mark.parametrize("bin_bytes", [1, 2])
def test_rebin_categorical_kernel_source(tmp_path, bin_bytes):
    """RebinCategoricalKernel = HistogramCuts::SearchCatBin on the categorical columns, other columns untouched."""
    exe = _build("emu_updater")
    rng = np.random.default_rng(5)
    n, nf = 700, 6
    ks = {1: 5, 4: 300 if bin_bytes == 2 else 40}       # categorical columns and their category counts
    x = rng.normal(size=(n, nf)).astype(np.float32)
    cut_values, cut_ptrs = [], [0]
    for f in range(nf):
        if f in ks:
            x[:, f] = rng.integers(0, ks[f] + 3, size=n).astype(np.float32) + rng.choice([0.0, 0.4], size=n)  # beyond the max too
            x[rng.random(n) < 0.1, f] = np.nan
            cut_values += [float(v) for v in range(ks[f])]
        else:
            cut_values += sorted(rng.normal(size=6).tolist())
        cut_ptrs.append(len(cut_values))
    cp, cv = np.array(cut_ptrs, dtype=np.uint32), np.array(cut_values, dtype=np.float32)
    row_bytes = 32 * bin_bytes
    null_symbol = int(np.diff(cp).max())
    dt = np.uint8 if bin_bytes == 1 else np.uint16
    page = rng.integers(0, 5, size=(n, 32)).astype(dt)                 # whatever the numerical kernel left there
    d = str(tmp_path)
    np.array([n, nf, nf, row_bytes, bin_bytes, null_symbol], dtype=np.uint32).tofile(os.path.join(d, "meta.bin"))
    x.tofile(os.path.join(d, "x.bin"))
    np.array(sorted(ks), dtype=np.uint32).tofile(os.path.join(d, "cat_features.bin"))
    cp.tofile(os.path.join(d, "cut_ptrs.bin"))
    cv.tofile(os.path.join(d, "cut_values.bin"))
    page.tofile(os.path.join(d, "bins.bin"))
    _run(exe, "rebin", d)
    out = np.fromfile(os.path.join(d, "bins_out.bin"), dtype=dt).reshape(n, 32)
    want = page.copy()
    for f, k in ks.items():
        col = x[:, f]
        v = np.trunc(col)
        b = np.minimum(np.searchsorted(cv[cp[f]:cp[f + 1]], v, side="left"), k - 1)
        want[:, f] = np.where(np.isnan(col), null_symbol, b).astype(dt)
    assert np.array_equal(out, want)


# ---------------------------------------------------------------------------------------------------------------------
# Gradient-based row sampling (csrc/sampling.cu)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("subsample,seed,obj", [(0.3, 12345, "sq"), (0.7, 4294967295, "sq"), (0.1, 7, "logistic"),
                                                (0.95, 0, "sq")])
def test_gradient_based_sampling_kernel_sources(tmp_path, subsample, seed, obj):
    """The kernels of csrc/sampling.cu around a CPU sort / integer scan: sampled pairs, threshold, the recomputed root
    sum and hessian flags against orc_gradient_based_sample.  Some rows carry a zero pair (they must stay untouched and
    unsampled)."""
    from tests.helpers import logistic_gpair
    exe = _build("emu_sampling")
    _, y = make_regression(5000, 4, seed=31)
    gp = squared_error_gpair(np.zeros_like(y), y) if obj == "sq" else logistic_gpair(y, (y > 0).astype(np.float32))
    gp[::97] = 0.0
    scale, _ = orc.quantiser(gp)
    gq = orc.quantise(gp, scale)
    d = str(tmp_path)
    gq.tofile(os.path.join(d, "gq.bin"))
    np.asarray(scale, dtype=np.float64).tofile(os.path.join(d, "quant.bin"))
    np.array([subsample], dtype=np.float32).tofile(os.path.join(d, "param.bin"))
    np.array([seed], dtype=np.uint32).tofile(os.path.join(d, "seed.bin"))
    subprocess.run([exe, d], check=True, timeout=900)
    out = np.fromfile(os.path.join(d, "gq_out.bin"), dtype=np.int64).reshape(-1, 2)
    mu = float(np.fromfile(os.path.join(d, "mu.bin"), dtype=np.float32)[0])
    rag = np.fromfile(os.path.join(d, "rag.bin"), dtype=np.float32)
    root = np.fromfile(os.path.join(d, "root.bin"), dtype=np.int64)
    want, want_mu, want_tidx, want_rag = orc.gradient_based_sample(gq, scale, subsample, seed)
    assert np.array_equal(rag.view(np.uint32), want_rag.view(np.uint32))
    assert np.float32(mu).view(np.uint32) == np.float32(want_mu).view(np.uint32)
    assert np.array_equal(out, want)
    assert (int(root[0]), int(root[1])) == (int(want[:, 0].sum()), int(want[:, 1].sum()))
    assert int(root[2]) == int(np.bitwise_or.reduce(want[:, 1].astype(np.uint64) & np.uint64(0xFFFFFFFF)))
    assert int(root[3]) == int((want[:, 1] == 0).any())


# ---------------------------------------------------------------------------------------------------------------------
# A whole tree: the launch sequence of UpdateTreeImpl with the real control / evaluation / partition / finalise kernels
# (tests/emu/emu_tree.cc); only the histogram build is a CPU loop over the kernels' own work lists.
# ---------------------------------------------------------------------------------------------------------------------
TREE_NODE = np.dtype([("parent", "<i4"), ("left", "<i4"), ("right", "<i4"), ("findex", "<i4"), ("fvalue", "<f4"),
                      ("default_left", "<i4"), ("is_leaf", "<i4"), ("leaf_value", "<f4"), ("loss_chg", "<f4"),
                      ("sum_hess", "<f4"), ("base_weight", "<f4"), ("depth", "<i4")])


def _emulate_tree(tmp_path, gb, cp, cv, gp, params, mono=None, is_cat=None, onehot=4, threshold=64, ic_sets=None,
                  node_sets=None, feature_set=None, n_levels=0, col_seeds=None, col_tree_set=None, bylevel=1.0, bynode=1.0,
                  sparse=False, hist_mode=0, page_rows=0, level_batch=0, cache_slots=0, world=1, full_exchange=False,
                  shard_ptr=None):
    exe = _build("emu_tree")
    n, nf = gb.shape
    max_bins = int(np.diff(cp).max())
    has_missing = bool((gb < 0).any())
    bin_bytes = 2 if max_bins + (1 if has_missing else 0) > 256 else 1
    null_symbol = max_bins if has_missing else 0xFFFFFFFF
    page = _page_bytes(gb, cp, bin_bytes, null_symbol if has_missing else 0)
    row_stride = 0
    if sparse:   # the reference's sparse ELLPACK: a row's present global bins in feature order, padded with total_bins
        row_stride = max(1, int((gb >= 0).sum(axis=1).max()))
        ell = np.full((n, row_stride), cv.size, dtype=np.uint32)
        for r in range(n):
            present = gb[r][gb[r] >= 0]
            ell[r, :present.size] = present
        page = ell.view(np.uint8).reshape(n, row_stride * 4)
        bin_bytes, null_symbol = 4, cv.size
    scale, _ = orc.quantiser(gp)
    gq = orc.quantise(gp, scale)
    cat_words = int(cv.max()) // 32 + 1
    max_cat_bins = int(max([cp[f + 1] - cp[f] for f in range(nf) if is_cat is not None and is_cat[f]] or [1]))
    d = str(tmp_path)
    p = orc.Param.make(**params)
    fs = None if feature_set is None else np.ascontiguousarray(feature_set, dtype=np.int32)
    col_meta = [0, 0, 0, 1, 1]
    if col_seeds is not None:
        ts_ = np.ascontiguousarray(col_tree_set if col_tree_set is not None else np.arange(nf), dtype=np.int32)
        n_level = ts_.size if bylevel == 1.0 else max(1, int(np.float32(bylevel) * np.float32(ts_.size)))
        n_node = n_level if bynode == 1.0 else max(1, int(np.float32(bynode) * np.float32(n_level)))
        col_meta = [len(col_seeds), n_level, n_node, 1 if bylevel == 1.0 else 0, 1 if bynode == 1.0 else 0]
        os.makedirs(str(tmp_path), exist_ok=True)
        np.ascontiguousarray(col_seeds, dtype=np.uint32).tofile(os.path.join(str(tmp_path), "col_seeds.bin"))
        ts_.tofile(os.path.join(str(tmp_path), "col_tree_set.bin"))
    np.array([n, nf, page.shape[1], bin_bytes, null_symbol, 0 if mono is None else 1, 0 if is_cat is None else 1, onehot,
              threshold, cat_words, (max_cat_bins + 31) // 32, max_cat_bins, 0 if ic_sets is None else len(ic_sets),
              0 if node_sets is None else node_sets.shape[0], 0 if node_sets is None else node_sets.shape[1],
              0 if fs is None else 1, 0 if fs is None else fs.shape[-1], n_levels] + col_meta + [row_stride, hist_mode, page_rows, level_batch, cache_slots,
                                                                                         world, int(full_exchange)],   # [29]: 0 slices, 1 fused allreduce, 2 feature blocks
             dtype=np.int64).tofile(os.path.join(d, "meta.bin"))
    np.array([p.learning_rate, p.reg_lambda, p.reg_alpha, p.min_child_weight, p.max_delta_step, p.min_split_loss],
             dtype=np.float32).tofile(os.path.join(d, "param_f.bin"))
    np.array([p.max_depth, p.max_leaves, p.grow_policy], dtype=np.int32).tofile(os.path.join(d, "param_i.bin"))
    np.asarray(scale, dtype=np.float64).tofile(os.path.join(d, "quant.bin"))
    np.array([gq[:, 0].sum(), gq[:, 1].sum(), int((gq[:, 1] == 0).any()),
              int(np.bitwise_or.reduce(gq[:, 1].astype(np.uint64) & np.uint64(0xFFFFFFFF)))],
             dtype=np.int64).tofile(os.path.join(d, "root.bin"))
    page.tofile(os.path.join(d, "bins.bin"))
    cp.astype(np.uint32).tofile(os.path.join(d, "cut_ptrs.bin"))
    cv.astype(np.float32).tofile(os.path.join(d, "cut_values.bin"))
    gq.tofile(os.path.join(d, "gq.bin"))
    if mono is not None:
        m = np.zeros(nf, dtype=np.int32)
        m[:len(mono)] = mono
        m.tofile(os.path.join(d, "mono.bin"))
    if is_cat is not None:
        np.asarray(is_cat, dtype=np.uint8).tofile(os.path.join(d, "is_cat.bin"))
    if ic_sets is not None:
        np.cumsum([0] + [len(s) for s in ic_sets]).astype(np.int32).tofile(os.path.join(d, "ic_set_ptr.bin"))
        np.array([f for s in ic_sets for f in s], dtype=np.int32).tofile(os.path.join(d, "ic_set_feat.bin"))
    if node_sets is not None:
        np.ascontiguousarray(node_sets, dtype=np.int32).tofile(os.path.join(d, "node_sets.bin"))
    if fs is not None:
        fs.tofile(os.path.join(d, "feature_set.bin"))
    if shard_ptr is not None:
        np.asarray(shard_ptr, dtype=np.int64).tofile(os.path.join(d, "shard_ptr.bin"))
    subprocess.run([exe, d], check=True, timeout=1800)
    if world > 1:   # one result set per rank (a process each, tests/emu/emu_tree.cc): rows are sharded, trees replicated
        return [(np.fromfile(os.path.join(d, "nodes.bin_r%d" % r), dtype=TREE_NODE),
                 np.fromfile(os.path.join(d, "position.bin_r%d" % r), dtype=np.int32)) for r in range(world)]
    nodes = np.fromfile(os.path.join(d, "nodes.bin"), dtype=TREE_NODE)
    pos = np.fromfile(os.path.join(d, "position.bin"), dtype=np.int32)
    flags = np.fromfile(os.path.join(d, "is_cat.bin"), dtype=np.int32)
    bits = np.fromfile(os.path.join(d, "cat_bits.bin"), dtype=np.uint32).reshape(-1, cat_words)
    counts = np.fromfile(os.path.join(d, "eval_count.bin"), dtype=np.int32)
    return nodes, pos, flags, bits, counts


def _same_tree(nodes, pos, o, flags=None, bits=None):
    assert len(nodes) == len(o["nodes"])
    for i, (a, b) in enumerate(zip(nodes, o["nodes"])):
        assert (int(a["parent"]), int(a["left"]), int(a["right"]), bool(a["is_leaf"]), int(a["depth"])) == \
            (b.parent, b.left, b.right, bool(b.is_leaf), b.depth), i
        if not b.is_leaf:
            assert (int(a["findex"]), bool(a["default_left"])) == (b.findex, bool(b.default_left)), i
            cat_node = "is_cat" in o and o["is_cat"][i]
            if cat_node:
                got = {32 * w + (31 - k) for w in range(bits.shape[1]) for k in range(32) if (int(bits[i, w]) >> k) & 1}
                assert flags[i] == 1 and np.isnan(a["fvalue"]) and got == o["cat_sets"][i], i
            else:
                assert np.float32(a["fvalue"]).view(np.uint32) == np.float32(b.fvalue).view(np.uint32), i
                assert flags is None or flags[i] == 0, i
            assert np.float32(a["loss_chg"]).view(np.uint32) == np.float32(b.loss_chg).view(np.uint32), i
        else:
            assert np.float32(a["leaf_value"]).view(np.uint32) == np.float32(b.leaf_value).view(np.uint32), i
        assert np.float32(a["sum_hess"]).view(np.uint32) == np.float32(b.sum_hess).view(np.uint32), i
        assert np.float32(a["base_weight"]).view(np.uint32) == np.float32(b.base_weight).view(np.uint32), i
    assert np.array_equal(pos, o["position"])


def _tree_data(n=2500, f=8, max_bin=24, missing=0.0, seed=13, zero_hess=False):
    x, y = make_regression(n, f, seed=seed, missing=missing)
    y = (y + np.nan_to_num(x[:, 0]) * np.nan_to_num(x[:, 1]) - np.nan_to_num(x[:, 2]) * np.nan_to_num(x[:, 3])).astype(np.float32)
    cp, cv = make_cuts(x, max_bin)
    gb = orc.bin_dense(x, cp, cv)
    gp = squared_error_gpair(np.zeros_like(y), y)
    if zero_hess:
        gp[::11, 1] = 0.0
    return gb, cp, cv, gp


@pytest.mark.parametrize("name,params,extra", [
    ("depthwise", dict(max_depth=5, eta=0.5), {}),
    ("lossguide", dict(max_depth=0, max_leaves=13, grow_policy="lossguide", eta=0.5), {}),
    ("depthwise_max_leaves", dict(max_depth=6, max_leaves=9, eta=0.5), {}),
    ("missing_u16_zero_hess", dict(max_depth=4, eta=0.3, min_child_weight=0.0), dict(missing=0.15, max_bin=300, zero_hess=True)),
    ("monotone", dict(max_depth=5, eta=0.5), dict(mono=[1, -1, 0, 1, 0, -1, 0, 0])),
    ("level_sets", dict(max_depth=5, eta=0.5), dict(level_sets=True)),
], ids=lambda v: v if isinstance(v, str) else "")
def test_whole_tree_kernel_sources(tmp_path, name, params, extra):
    gb, cp, cv, gp = _tree_data(missing=extra.get("missing", 0.0), max_bin=extra.get("max_bin", 24),
                                zero_hess=extra.get("zero_hess", False))
    kw, okw = {}, {}
    if extra.get("mono"):
        kw["mono"] = okw["monotone"] = extra["mono"]
    if extra.get("level_sets"):
        rng = np.random.default_rng(4)
        levels = np.stack([rng.permutation(8)[:5] for _ in range(5)]).astype(np.int32)   # unsorted: position = tie rank
        kw.update(feature_set=levels, n_levels=5)
        okw["level_feature_sets"] = levels
    nodes, pos, _, _, _ = _emulate_tree(tmp_path, gb, cp, cv, gp, params, **kw)
    _same_tree(nodes, pos, orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), **okw))


@pytest.mark.parametrize("params", [dict(max_depth=5, eta=0.5), dict(max_depth=0, max_leaves=12, grow_policy="lossguide", eta=0.5)],
                         ids=["depthwise", "lossguide"])
def test_whole_tree_kernel_sources_colsample_bynode(tmp_path, params):
    """The k-th node evaluated takes the k-th draw: NodeSetLevelKernel + the evaluation order of the real control
    kernels against the oracle's evaluation order, and the number of draws used."""
    gb, cp, cv, gp = _tree_data(seed=14)
    p = orc.Param.make(**params)
    max_nodes = min((1 << (p.max_depth + 1)) - 1 if p.max_depth > 0 else 1 << 30, 2 * p.max_leaves - 1 if p.max_leaves > 0 else 1 << 30)
    rng = np.random.default_rng(15)
    sets = np.stack([np.sort(rng.permutation(8)[:3]) for _ in range(max_nodes)]).astype(np.int32)
    nodes, pos, _, _, counts = _emulate_tree(tmp_path, gb, cp, cv, gp, params, node_sets=sets)
    o = orc.update_tree(gb, cp, cv, gp, p, node_feature_sets=sets)
    _same_tree(nodes, pos, o)
    assert int(counts[0]) == o["eval_count"]


@pytest.mark.parametrize("with_levels", [False, True], ids=["all_features", "level_sets"])
def test_whole_tree_kernel_sources_interaction_constraints(tmp_path, with_levels):
    """Overlapping sets: the tree must follow the shared Query buffer (the reference's behaviour), not the per-query
    answers -- IcLevelKernel and the constrained evaluation kernels chained over the levels of a real tree."""
    from tests.helpers import make_interaction_data
    x, y = make_interaction_data(dict(seed=201, n=2500, f=6))
    cp, cv = make_cuts(x, 24)
    gb = orc.bin_dense(x, cp, cv)
    gp = squared_error_gpair(np.zeros_like(y), y)
    chain = [[0, 1], [1, 2], [2, 3], [3, 4], [4, 5]]
    params = dict(max_depth=5, eta=0.5)
    kw, okw = {}, {}
    if with_levels:
        rng = np.random.default_rng(6)
        levels = np.stack([np.sort(rng.permutation(6)[:5]) for _ in range(5)]).astype(np.int32)
        kw.update(feature_set=levels, n_levels=5)
        okw["level_feature_sets"] = levels
    nodes, pos, _, _, _ = _emulate_tree(tmp_path, gb, cp, cv, gp, params, ic_sets=chain, **kw)
    o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), interaction=chain, **okw)
    _same_tree(nodes, pos, o)
    per_query = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), interaction=chain, interaction_shared_buffer=False, **okw)
    if not with_levels:
        assert [nd.findex for nd in per_query["nodes"]] != [int(v) for v in nodes["findex"]]


@pytest.mark.parametrize("onehot,params", [(4, dict(max_depth=5, eta=0.5)), (100, dict(max_depth=4, eta=0.5)),
                                           (1, dict(max_depth=0, max_leaves=10, grow_policy="lossguide", eta=0.5))],
                         ids=["mixed_depthwise", "all_onehot", "all_partition_lossguide"])
def test_whole_tree_kernel_sources_categorical(tmp_path, onehot, params):
    """Categorical splits through ApplyKernel<true>, PartitionCountKernel<., true> and FinaliseWindowKernel<., true>:
    category sets, NaN conditions and the position of every row."""
    from tests.helpers import make_categorical_data
    from tests.test_reference_golden import categorical_global_bins
    case = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_cpu_categorical_trees.json")))["cases"][0]
    x, y, types = make_categorical_data(dict(case, n=2500))
    is_cat = [1 if t == "c" else 0 for t in types]
    cp = np.array(case["cut_ptrs"], dtype=np.uint32)
    cv = np.array(case["cut_values_bits"], dtype=np.uint32).view(np.float32)
    gb = categorical_global_bins(x, cp, cv, is_cat)
    gp = squared_error_gpair(np.zeros_like(y), y)
    nodes, pos, flags, bits, _ = _emulate_tree(tmp_path, gb, cp, cv, gp, params, is_cat=is_cat, onehot=onehot)
    o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), feature_is_cat=is_cat, max_cat_to_onehot=onehot)
    _same_tree(nodes, pos, o, flags, bits)
    assert int(o["is_cat"].sum()) >= 3


def test_whole_tree_kernel_sources_combined(tmp_path):
    """Categorical features + colsample_bynode + monotone constraints on the numerical features in one tree."""
    from tests.helpers import make_categorical_data
    from tests.test_reference_golden import categorical_global_bins
    case = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_cpu_categorical_trees.json")))["cases"][0]
    x, y, types = make_categorical_data(dict(case, n=2500, seed=77))
    is_cat = [1 if t == "c" else 0 for t in types]
    cp = np.array(case["cut_ptrs"], dtype=np.uint32)
    cv = np.array(case["cut_values_bits"], dtype=np.uint32).view(np.float32)
    gb = categorical_global_bins(x, cp, cv, is_cat)
    gp = squared_error_gpair(np.zeros_like(y), y)
    params = dict(max_depth=5, eta=0.5)
    mono = [0, 0, 0, 0, 1, -1, 0]
    rng = np.random.default_rng(21)
    sets = np.stack([np.sort(rng.permutation(7)[:4]) for _ in range(63)]).astype(np.int32)
    nodes, pos, flags, bits, counts = _emulate_tree(tmp_path, gb, cp, cv, gp, params, is_cat=is_cat, node_sets=sets, mono=mono)
    o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), feature_is_cat=is_cat, node_feature_sets=sets, monotone=mono)
    _same_tree(nodes, pos, o, flags, bits)
    assert int(counts[0]) == o["eval_count"] and int(o["is_cat"].sum()) >= 2


@pytest.mark.parametrize("bylevel,bynode,params", [
    (1.0, 0.5, dict(max_depth=5, eta=0.5)),
    (0.6, 0.5, dict(max_depth=5, eta=0.5)),
    (0.6, 1.0, dict(max_depth=5, eta=0.5)),
    (0.7, 0.6, dict(max_depth=0, max_leaves=12, grow_policy="lossguide", eta=0.5)),
], ids=["bynode", "bylevel_bynode", "bylevel", "bylevel_bynode_lossguide"])
def test_whole_tree_kernel_sources_column_sampler(tmp_path, bylevel, bynode, params):
    """The seed-driven column sampler (LevelDrawKernel / NodeDrawKernel, csrc/colsample.cu) inside a real tree: level
    sets drawn the first time a depth is evaluated, node sets per evaluated node, all from one stream of engine outputs
    whose cursor lives on the device -- incl. loss-guided growth, where depths come in data-dependent order -- against
    the oracle's lazy restatement, and the number of outputs consumed."""
    gb, cp, cv, gp = _tree_data(seed=16, f=10)
    rng = np.random.default_rng(17)
    seeds = [int(v) for v in rng.integers(0, 2 ** 32, size=200)]
    tree_set = np.sort(rng.permutation(10)[:8]).astype(np.int32)       # a colsample_bytree draw
    nodes, pos, _, _, counts = _emulate_tree(tmp_path, gb, cp, cv, gp, params, col_seeds=seeds, col_tree_set=tree_set,
                                             bylevel=bylevel, bynode=bynode)
    o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), column_seeds=seeds, tree_set=tree_set,
                        colsample_bylevel=bylevel, colsample_bynode=bynode)
    _same_tree(nodes, pos, o)
    assert int(counts[2]) == o["draws_used"]
    for nd in o["nodes"]:
        if not nd.is_leaf:
            assert nd.findex in tree_set


@pytest.mark.parametrize("bylevel,bynode", [(1.0, 0.6), (0.8, 0.7), (0.7, 1.0)], ids=["bynode", "bylevel_bynode", "bylevel"])
def test_whole_tree_kernel_sources_column_sampler_with_interaction_constraints(tmp_path, bylevel, bynode):
    """Query(the node's own sampled list, nidx): IcLevelKernel with per-slot lists from NodeDrawKernel (or the lazily
    drawn level sets), overlapping constraint sets, the shared result buffer."""
    from tests.helpers import make_interaction_data
    x, y = make_interaction_data(dict(seed=202, n=2500, f=6))
    cp, cv = make_cuts(x, 24)
    gb = orc.bin_dense(x, cp, cv)
    gp = squared_error_gpair(np.zeros_like(y), y)
    chain = [[0, 1], [1, 2], [2, 3], [3, 4], [4, 5]]
    params = dict(max_depth=5, eta=0.5)
    rng = np.random.default_rng(19)
    seeds = [int(v) for v in rng.integers(0, 2 ** 32, size=200)]
    nodes, pos, _, _, counts = _emulate_tree(tmp_path, gb, cp, cv, gp, params, ic_sets=chain, col_seeds=seeds,
                                             bylevel=bylevel, bynode=bynode)
    o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), interaction=chain, column_seeds=seeds,
                        colsample_bylevel=bylevel, colsample_bynode=bynode)
    _same_tree(nodes, pos, o)
    assert int(counts[2]) == o["draws_used"] and len(o["nodes"]) > 7


@pytest.mark.parametrize("categorical", [False, True], ids=["numerical", "categorical"])
def test_whole_tree_kernel_sources_sparse_page(tmp_path, categorical):
    """The sparse ELLPACK layout (global-bin symbols, row searched by feature) through partition and FinalisePosition,
    with and without categorical routing."""
    if categorical:
        from tests.helpers import make_categorical_data
        from tests.test_reference_golden import categorical_global_bins
        case = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_cpu_categorical_trees.json")))["cases"][0]
        x, y, types = make_categorical_data(dict(case, n=2000, seed=78))
        rng = np.random.default_rng(3)
        x[rng.random(x.shape) < 0.5] = np.nan                       # a sparse matrix
        is_cat = [1 if t == "c" else 0 for t in types]
        cp = np.array(case["cut_ptrs"], dtype=np.uint32)
        cv = np.array(case["cut_values_bits"], dtype=np.uint32).view(np.float32)
        gb = categorical_global_bins(x, cp, cv, is_cat)
        kw, okw = dict(is_cat=is_cat), dict(feature_is_cat=is_cat)
    else:
        x, y = make_regression(2000, 12, seed=41, missing=0.6)
        cp, cv = make_cuts(x, 24)
        gb = orc.bin_dense(x, cp, cv)
        kw, okw = {}, {}
    gp = squared_error_gpair(np.zeros_like(y), y)
    params = dict(max_depth=5, eta=0.5)
    nodes, pos, flags, bits, _ = _emulate_tree(tmp_path, gb, cp, cv, gp, params, sparse=True, **kw)
    o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), **okw)
    _same_tree(nodes, pos, o, flags if categorical else None, bits if categorical else None)


@pytest.mark.parametrize("slots,hist_mode,params", [
    (9, 0, dict(max_depth=6, eta=0.5)),                                          # windows of up to 32 slots: the floor
    (40, 1, dict(max_depth=6, eta=0.5)),                                         # ring of 39: the last levels wrap
    (7, 2, dict(max_depth=0, max_leaves=24, grow_policy="lossguide", eta=0.5)),  # loss-guide: old parents get recycled
    (12, 0, dict(max_depth=5, max_leaves=14, eta=0.5)),
], ids=["floor", "wrap_real_kernels", "lossguide_pair_kernel", "leaf_budget"])
def test_whole_tree_with_a_histogram_cache_smaller_than_the_tree(tmp_path, slots, hist_mode, params):
    """DeviceHistogramStorage's overflow (histogram.cuh:72-168, updater_gpu_hist.cu:309-325): with a cache of `slots`
    histograms the steps' slot windows wrap around a ring, ScanKernel finds parents whose histogram has been overwritten
    (slot ownership) and builds BOTH children there instead of subtracting.  Integer sums: the tree must not change."""
    n, f = 2500, 64 if hist_mode == 2 else 9
    x, y = make_regression(n, f, seed=43)
    cp, cv = make_cuts(x, 16)
    gb = orc.bin_dense(x, cp, cv)
    gp = squared_error_gpair(np.zeros_like(y), y)
    nodes, pos, _, _, _ = _emulate_tree(tmp_path, gb, cp, cv, gp, params, hist_mode=hist_mode, cache_slots=slots)
    _same_tree(nodes, pos, orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params)))


@pytest.mark.parametrize("batch,params", [
    (3, dict(max_depth=5, eta=0.5)),                                    # levels of 4, 8, 16 nodes in steps of 3
    (1, dict(max_depth=4, eta=0.5)),                                    # one node per step
    (5, dict(max_depth=6, max_leaves=19, eta=0.5)),                     # the leaf budget ends a level half way
    (2, dict(max_depth=5, eta=0.5, min_child_weight=40.0, min_split_loss=2.0)),   # invalid entries inside a level
    (4, dict(max_depth=0, max_leaves=13, eta=0.5)),                     # depth-wise under a leaf budget alone
], ids=["steps_of_3", "steps_of_1", "leaf_budget", "invalid_entries", "leaves_only"])
def test_whole_tree_with_node_batching(tmp_path, batch, params):
    """Node batching (Driver::Pop hands out at most max_node_batch_size entries of a level, driver.h:84-99;
    kMaxNodeBatchSize, updater_gpu_hist.cuh:34): a level is expanded by several steps of ApplyKernel / partition /
    build / evaluate.  The oracle restates the Driver with the same batch size; without a leaf budget the tree also
    equals the unbatched one."""
    n, f = 2500, 9
    x, y = make_regression(n, f, seed=41)
    cp, cv = make_cuts(x, 24)
    gb = orc.bin_dense(x, cp, cv)
    gp = squared_error_gpair(np.zeros_like(y), y)
    nodes, pos, _, _, _ = _emulate_tree(tmp_path, gb, cp, cv, gp, params, level_batch=batch)
    o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), node_batch=batch)
    _same_tree(nodes, pos, o)
    if not params.get("max_leaves"):
        _same_tree(nodes, pos, orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params)))


@pytest.mark.parametrize("name,hist_mode,page_rows,kw", [
    ("pair_u8_96f_pages_of_400", 2, 400, dict(f=96, max_bin=16)),
    ("banked_u8_missing_37f_pages_of_311", 1, 311, dict(f=37, max_bin=24, missing=0.1)),
    ("pair_fallback_logistic_64f_pages_of_1000", 2, 1000, dict(f=64, max_bin=16, logistic=True)),
    ("banked_u16_300bins_pages_of_64", 1, 64, dict(f=9, max_bin=300, missing=0.1)),
    ("generic_u16_600bins_pages_of_700", 1, 700, dict(f=6, max_bin=600)),
    ("cpu_hist_lossguide_pages_of_512", 0, 512, dict(f=12, max_bin=24, lossguide=True)),
], ids=lambda v: v if isinstance(v, str) else "")
def test_whole_tree_over_external_memory_pages(tmp_path, name, hist_mode, page_rows, kw):
    """External memory (b200hist_ellpack_create_paged, UpdateTreeImpl's for_each_page): every phase that reads the matrix
    -- the real histogram kernels with the page clip compiled in, PartitionCountKernel accumulating decision bits over
    the pages, FinaliseWindowKernel page by page -- sees ONE page at a time (an exact-size copy: reading a row outside it
    is out of bounds).  The tree, every position and the row partition must be those of the in-core matrix."""
    from tests.helpers import logistic_gpair
    n = 1500
    x, y = make_regression(n, kw["f"], seed=29, missing=kw.get("missing", 0.0))
    cp, cv = make_cuts(x, kw["max_bin"])
    gb = orc.bin_dense(x, cp, cv)
    if kw.get("logistic"):
        gp = logistic_gpair(np.zeros_like(y), (y > 0).astype(np.float32))
    else:
        gp = squared_error_gpair(np.zeros_like(y), y)
    params = dict(max_depth=4, eta=0.5)
    if kw.get("lossguide"):
        params = dict(max_depth=0, max_leaves=11, grow_policy="lossguide", eta=0.5)
    nodes, pos, _, _, _ = _emulate_tree(tmp_path, gb, cp, cv, gp, params, hist_mode=hist_mode, page_rows=page_rows)
    _same_tree(nodes, pos, orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params)))


@pytest.mark.parametrize("name,hist_mode,kw", [
    ("banked_u8_64f", 1, dict(f=64, max_bin=24)),
    ("banked_u8_missing_37f", 1, dict(f=37, max_bin=24, missing=0.1)),
    ("pair_u8_96f", 2, dict(f=96, max_bin=16)),
    ("pair_fallback_logistic_64f", 2, dict(f=64, max_bin=16, logistic=True)),
    ("banked_u16_300bins", 1, dict(f=9, max_bin=300, missing=0.1)),
    ("generic_u16_600bins", 1, dict(f=6, max_bin=600)),
    ("sparse", 1, dict(f=12, max_bin=24, missing=0.6, sparse=True)),
], ids=lambda v: v if isinstance(v, str) else "")
def test_whole_tree_with_the_real_histogram_kernels(tmp_path, name, hist_mode, kw):
    """The hot kernels themselves under the emulator (their inline-PTX shared atomics replaced by emulator twins, the
    rest of the text unchanged): bank-mapped 32-feature kernel on uint8 and uint16 symbols, the 64-feature pair kernel
    and its fallback when hessians carry low words, the generic kernel beyond ~440 symbols, the sparse kernel --
    chosen like LaunchBuildHist chooses them, inside a whole tree, so that every node histogram feeds evaluation."""
    from tests.helpers import logistic_gpair
    n = 1500
    x, y = make_regression(n, kw["f"], seed=23, missing=kw.get("missing", 0.0))
    cp, cv = make_cuts(x, kw["max_bin"])
    gb = orc.bin_dense(x, cp, cv)
    if kw.get("logistic"):
        gp = logistic_gpair(np.zeros_like(y), (y > 0).astype(np.float32))
    else:
        gp = squared_error_gpair(np.zeros_like(y), y)
    params = dict(max_depth=4, eta=0.5)
    nodes, pos, _, _, _ = _emulate_tree(tmp_path, gb, cp, cv, gp, params, sparse=kw.get("sparse", False), hist_mode=hist_mode)
    _same_tree(nodes, pos, orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params)))


# Row split over several emulated GPUs: one PROCESS per rank, the histogram caches / flag arrays / scratch tables in a
# shared mapping (what IPC-mapped peer memory is on the device), the exchange kernels of comm.cu and their flag barriers
# running as written.  Every rank must grow the single-GPU tree and leave the oracle's positions for its rows.
@pytest.mark.parametrize("world,full_exchange,f,params,extra", [
    (2, False, 8, dict(max_depth=5, eta=0.5), {}),
    (3, False, 8, dict(max_depth=4, eta=0.5), dict(uneven=True)),
    (4, False, 10, dict(max_depth=5, eta=0.5), dict(missing=0.1)),
    (8, False, 8, dict(max_depth=4, eta=0.5), {}),
    (4, False, 9, dict(max_depth=0, max_leaves=11, grow_policy="lossguide", eta=0.5), {}),
    (4, False, 8, dict(max_depth=4, eta=0.5), dict(level_sets=True)),
    (3, False, 8, dict(max_depth=4, eta=0.5), dict(mono=[1, -1, 0, 1, 0, -1, 0, 0])),
    (4, False, 8, dict(max_depth=5, eta=0.5), dict(hist_mode=2)),
    (4, True, 8, dict(max_depth=4, eta=0.5), {}),
    (4, False, 8, dict(max_depth=6, eta=0.5), dict(cache_slots=20)),      # ring cache: lost parents -> both children built and exchanged
    (2, True, 8, dict(max_depth=6, eta=0.5), dict(cache_slots=20)),
    (3, False, 8, dict(max_depth=6, eta=0.5), dict(level_batch=5)),       # levels expanded in node batches
    (5, False, 4, dict(max_depth=4, eta=0.5), {}),        # fewer features than ranks: the fused allreduce serves
], ids=lambda v: str(v) if isinstance(v, (int, bool)) else "")
def test_whole_tree_row_split_over_emulated_ranks(tmp_path, world, full_exchange, f, params, extra):
    gb, cp, cv, gp = _tree_data(n=2100, f=f, missing=extra.get("missing", 0.0), seed=29)
    n = gb.shape[0]
    kw, okw = {}, {}
    if extra.get("mono"):
        kw["mono"] = okw["monotone"] = extra["mono"]
    if extra.get("level_sets"):
        rng = np.random.default_rng(5)
        levels = np.stack([rng.permutation(f)[:5] for _ in range(4)]).astype(np.int32)
        kw.update(feature_set=levels, n_levels=4)
        okw["level_feature_sets"] = levels
    shard_ptr = None
    if extra.get("uneven"):
        shard_ptr = [0, 100, 1700, n]
    ranks = _emulate_tree(tmp_path, gb, cp, cv, gp, params, world=world, full_exchange=full_exchange, shard_ptr=shard_ptr,
                          hist_mode=extra.get("hist_mode", 0), cache_slots=extra.get("cache_slots", 0),
                          level_batch=extra.get("level_batch", 0), **kw)
    if extra.get("level_batch"):
        okw["node_batch"] = extra["level_batch"]
    o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params), **okw)   # which child is built does not show in exact sums
    bounds = shard_ptr if shard_ptr is not None else [n * r // world for r in range(world + 1)]
    for r, (nodes, pos) in enumerate(ranks):
        _same_tree(nodes, pos, dict(o, position=o["position"][bounds[r]:bounds[r + 1]]))


@pytest.mark.parametrize("world,what", [(4, "node_sets"), (3, "interaction"), (4, "column_sampler"),
                                        (2, "column_sampler_with_interaction")])
def test_row_split_over_emulated_ranks_with_per_node_feature_lists(tmp_path, world, what):
    """Evaluation by feature slice when every evaluation slot has its OWN feature list (colsample_bynode draws, the
    interaction-constraint query result): a rank scans the list entries whose feature it owns, and ScratchExchangeKernel
    must move exactly those entries -- both look a slot's feature up through the same list."""
    from tests.helpers import make_interaction_data
    params = dict(max_depth=5, eta=0.5)
    p = orc.Param.make(**params)
    kw, okw = {}, {}
    if what == "node_sets":
        gb, cp, cv, gp = _tree_data(n=2100, seed=31)
        rng = np.random.default_rng(15)
        sets = np.stack([np.sort(rng.permutation(8)[:3]) for _ in range((1 << (p.max_depth + 1)) - 1)]).astype(np.int32)
        kw["node_sets"], okw["node_feature_sets"] = sets, sets
    else:
        x, y = make_interaction_data(dict(seed=203, n=2100, f=6))
        cp, cv = make_cuts(x, 24)
        gb = orc.bin_dense(x, cp, cv)
        gp = squared_error_gpair(np.zeros_like(y), y)
        if "interaction" in what:
            chain = [[0, 1], [1, 2], [2, 3], [3, 4], [4, 5]]
            kw["ic_sets"], okw["interaction"] = chain, chain
        if "column_sampler" in what:
            rng = np.random.default_rng(18)
            seeds = [int(v) for v in rng.integers(0, 2 ** 32, size=200)]
            kw.update(col_seeds=seeds, bylevel=0.8, bynode=0.7)
            okw.update(column_seeds=seeds, colsample_bylevel=0.8, colsample_bynode=0.7)
    n = gb.shape[0]
    ranks = _emulate_tree(tmp_path, gb, cp, cv, gp, params, world=world, **kw)
    o = orc.update_tree(gb, cp, cv, gp, p, **okw)
    for r, (nodes, pos) in enumerate(ranks):
        _same_tree(nodes, pos, dict(o, position=o["position"][n * r // world:n * (r + 1) // world]))


@pytest.mark.parametrize("world,f,params", [(2, 8, dict(max_depth=5, eta=0.5)), (4, 10, dict(max_depth=4, eta=0.5)),
                                            (3, 8, dict(max_depth=0, max_leaves=10, grow_policy="lossguide", eta=0.5)),
                                            (8, 8, dict(max_depth=4, eta=0.5))])
def test_feature_block_shard_over_emulated_ranks(tmp_path, world, f, params):
    """north_star's decomposition: all rows on every rank, a contiguous feature block each; HistExchangeKernel pushes the
    block's bin range of the level's built histograms into every peer, everything else is replicated."""
    gb, cp, cv, gp = _tree_data(n=1500, f=f, missing=0.05, seed=37)
    ranks = _emulate_tree(tmp_path, gb, cp, cv, gp, params, world=world, full_exchange=2)
    o = orc.update_tree(gb, cp, cv, gp, orc.Param.make(**params))
    for nodes, pos in ranks:
        _same_tree(nodes, pos, o)
