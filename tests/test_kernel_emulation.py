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


def _build(name):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, name)
    srcs = [os.path.join(EMU, name + ".cc"), os.path.join(EMU, "cuda_emu.cc")]
    csrc = os.path.join(ROOT, "xgboost_b200", "csrc")
    deps = srcs + [os.path.join(EMU, "cuda_emu.h")] + \
        [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-pthread", "-I" + os.path.join(EMU, "include"), "-I" + EMU,
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
                    lbin = int(np.searchsorted(cv[beg:end], np.float32(nd.fvalue), side="right"))
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


@pytest.mark.parametrize("kind", ["numerical_u8", "missing_u16", "categorical"])
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
        x, y = make_regression(3000, 9, seed=5, missing=0.1 if kind == "missing_u16" else 0.0)
        cp, cv = make_cuts(x, 300 if kind == "missing_u16" else 32)
        gb = orc.bin_dense(x, cp, cv)
        kw = {}
    pred = np.zeros(len(y), dtype=np.float32)
    trees = []
    for _ in range(4):
        o = orc.update_tree(gb, cp, cv, squared_error_gpair(pred, y), orc.Param.make(max_depth=4, eta=0.5), **kw)
        trees.append(o)
        pred = orc.update_prediction_cache(o["nodes"], o["position"], pred)
    max_bins = int(np.diff(cp).max())
    has_missing = bool((gb < 0).any())
    bin_bytes = 2 if max_bins + (1 if has_missing else 0) > 256 else 1
    null_symbol = max_bins if has_missing else 0xFFFFFFFF
    bin_words = (max_bins + 31) // 32
    nodes, roots, bits = _forest_arrays(trees, cp, cv, bin_words)
    page = _page_bytes(gb, cp, bin_bytes, null_symbol if has_missing else 0)
    base = np.full(len(y), 0.25, dtype=np.float32)
    d = str(tmp_path)
    np.array([len(y), gb.shape[1], page.shape[1], bin_bytes, null_symbol, len(trees), bin_words],
             dtype=np.uint32).tofile(os.path.join(d, "meta.bin"))
    page.tofile(os.path.join(d, "bins.bin"))
    nodes.tofile(os.path.join(d, "nodes.bin"))
    roots.tofile(os.path.join(d, "roots.bin"))
    bits.tofile(os.path.join(d, "bin_bits.bin"))
    cp.tofile(os.path.join(d, "cut_ptrs.bin"))
    cv.tofile(os.path.join(d, "cut_values.bin"))
    base.tofile(os.path.join(d, "base.bin"))
    _run(exe, "predict", d)
    margin = np.fromfile(os.path.join(d, "margin.bin"), dtype=np.float32)
    leaves = np.fromfile(os.path.join(d, "leaves.bin"), dtype=np.int32).reshape(len(y), len(trees))
    want = orc.predict_margin(gb, cp, cv, trees, base=base)
    assert np.array_equal(margin.view(np.uint32), want.view(np.uint32))
    for t, o in enumerate(trees):
        assert np.array_equal(leaves[:, t], orc.predict_leaf(gb, cp, cv, o["nodes"], o.get("is_cat"), o.get("cat_sets")))
