"""ctypes wrapper of oracle/hist_oracle.c -- the CPU restatement of the reference `gpu_hist`
algorithm.  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by xgboost_b200/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, byref, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "hist_oracle.c")
_LIB = os.path.join(_HERE, "liboracle.so")


class Param(Structure):
    _fields_ = [("max_depth", c_int32), ("max_leaves", c_int32), ("grow_policy", c_int32),
                ("learning_rate", c_float), ("reg_lambda", c_float), ("reg_alpha", c_float),
                ("min_child_weight", c_float), ("max_delta_step", c_float), ("min_split_loss", c_float)]

    @classmethod
    def make(cls, **kw):
        p = cls(6, 0, 0, 0.3, 1.0, 0.0, 1.0, 0.0, 0.0)
        alias = {"eta": "learning_rate", "lambda": "reg_lambda", "alpha": "reg_alpha", "gamma": "min_split_loss"}
        for k, v in kw.items():
            k = alias.get(k, k)
            if k == "grow_policy" and isinstance(v, str):
                v = {"depthwise": 0, "lossguide": 1}[v]
            setattr(p, k, v)
        return p


class GPair64(Structure):
    _fields_ = [("g", c_int64), ("h", c_int64)]


class Candidate(Structure):
    _fields_ = [("loss_chg", c_float), ("dir", c_int32), ("findex", c_int32), ("fvalue", c_float),
                ("left_sum", GPair64), ("right_sum", GPair64),
                ("base_weight", c_float), ("left_weight", c_float), ("right_weight", c_float),
                ("split_gbin", c_int32)]


class Node(Structure):
    _fields_ = [("parent", c_int32), ("left", c_int32), ("right", c_int32), ("findex", c_int32),
                ("fvalue", c_float), ("default_left", c_int32), ("is_leaf", c_int32), ("leaf_value", c_float),
                ("loss_chg", c_float), ("sum_hess", c_float), ("base_weight", c_float), ("depth", c_int32)]


_lib = None


def build() -> str:
    if (not os.path.exists(_LIB)) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", _LIB, _SRC,
                        "-lm"], check=True)
    return _LIB


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.orc_search_bin.restype = c_int32
        _lib.orc_search_bin.argtypes = [c_float, c_void_p, c_uint32, c_uint32]
        _lib.orc_symbol_bits.restype = c_uint32
        _lib.orc_symbol_bits.argtypes = [c_uint64]
        _lib.orc_ellpack_buffer_size.restype = c_uint64
        _lib.orc_ellpack_buffer_size.argtypes = [c_uint64, c_uint64]
        _lib.orc_ellpack_read.restype = c_uint32
        _lib.orc_ellpack_read.argtypes = [c_void_p, c_uint64, c_uint64]
        _lib.orc_add64_as32.restype = c_int64
        _lib.orc_add64_as32.argtypes = [c_int64, c_int64]
        _lib.orc_partition.restype = c_uint64
        _lib.orc_update_tree.restype = c_int32
    return _lib


def _p(a: np.ndarray):
    return c_void_p(a.ctypes.data)


def bin_dense(x: np.ndarray, cut_ptrs: np.ndarray, cut_values: np.ndarray) -> np.ndarray:
    """Global bin index per entry, -1 for NaN (orc_bin_dense)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, f = x.shape
    cp = np.ascontiguousarray(cut_ptrs, dtype=np.uint32)
    cv = np.ascontiguousarray(cut_values, dtype=np.float32)
    out = np.empty((n, f), dtype=np.int32)
    lib().orc_bin_dense(_p(x), c_uint64(n), c_uint32(f), _p(cp), _p(cv), _p(out))
    return out


def ellpack_pack(symbols: np.ndarray, n_symbols: int) -> np.ndarray:
    s = np.ascontiguousarray(symbols, dtype=np.uint32).ravel()
    size = lib().orc_ellpack_buffer_size(s.size, n_symbols)
    buf = np.zeros(size, dtype=np.uint8)
    lib().orc_ellpack_pack(_p(s), c_uint64(s.size), c_uint64(n_symbols), _p(buf))
    return buf


def ellpack_read(buf: np.ndarray, n_symbols: int, n_elements: int) -> np.ndarray:
    l = lib()
    return np.array([l.orc_ellpack_read(_p(buf), n_symbols, i) for i in range(n_elements)], dtype=np.uint32)


def quantiser(gpair: np.ndarray):
    """returns (scale[4] = to_fixed_g, to_fixed_h, to_float_g, to_float_h; sums[4])"""
    g = np.ascontiguousarray(gpair, dtype=np.float32)
    scale = (c_double * 4)()
    sums = (c_double * 4)()
    lib().orc_quantiser(_p(g), c_uint64(g.shape[0]), scale, sums)
    return np.array(scale[:], dtype=np.float64), np.array(sums[:], dtype=np.float64)


def quantise(gpair: np.ndarray, scale: np.ndarray) -> np.ndarray:
    g = np.ascontiguousarray(gpair, dtype=np.float32)
    sc = (c_double * 4)(*[float(v) for v in scale])
    out = np.empty((g.shape[0], 2), dtype=np.int64)
    lib().orc_quantise(_p(g), c_uint64(g.shape[0]), sc, _p(out))
    return out


def build_hist(gbins: np.ndarray, gq: np.ndarray, ridx, total_bins: int) -> np.ndarray:
    gb = np.ascontiguousarray(gbins, dtype=np.int32)
    q = np.ascontiguousarray(gq, dtype=np.int64)
    hist = np.zeros((total_bins, 2), dtype=np.int64)
    if ridx is None:
        lib().orc_build_hist(_p(gb), c_uint32(gb.shape[1]), _p(q), c_void_p(), c_uint64(gb.shape[0]), _p(hist))
    else:
        r = np.ascontiguousarray(ridx, dtype=np.uint32)
        lib().orc_build_hist(_p(gb), c_uint32(gb.shape[1]), _p(q), _p(r), c_uint64(r.size), _p(hist))
    return hist


def eval_node(hist: np.ndarray, parent_sum, cut_ptrs, cut_values, param: Param, scale, feature_set=None,
              want_gains: bool = False):
    h = np.ascontiguousarray(hist, dtype=np.int64)
    cp = np.ascontiguousarray(cut_ptrs, dtype=np.uint32)
    cv = np.ascontiguousarray(cut_values, dtype=np.float32)
    sc = (c_double * 4)(*[float(v) for v in scale])
    out = Candidate()
    gains = np.full((cv.size, 2), -np.inf, dtype=np.float32) if want_gains else None
    if feature_set is not None:
        fs = np.ascontiguousarray(feature_set, dtype=np.int32)
        fsp, nfs = _p(fs), fs.size
    else:
        fsp, nfs = c_void_p(), 0
    lib().orc_eval_node(_p(h), GPair64(int(parent_sum[0]), int(parent_sum[1])), _p(cp), _p(cv),
                        c_uint32(cp.size - 1), fsp, c_int32(nfs), byref(param), sc, byref(out),
                        _p(gains) if want_gains else c_void_p())
    return (out, gains) if want_gains else out


def eval_node_mono(hist: np.ndarray, parent_sum, cut_ptrs, cut_values, param: Param, scale, feature_set=None,
                   monotone=None, lower=-3.402823466e+38, upper=3.402823466e+38):
    """orc_eval_node_mono: eval_node under monotone constraints (sign per feature, weight bounds of this node)."""
    h = np.ascontiguousarray(hist, dtype=np.int64)
    cp = np.ascontiguousarray(cut_ptrs, dtype=np.uint32)
    cv = np.ascontiguousarray(cut_values, dtype=np.float32)
    sc = (c_double * 4)(*[float(v) for v in scale])
    out = Candidate()
    if feature_set is not None:
        fs = np.ascontiguousarray(feature_set, dtype=np.int32)
        fsp, nfs = _p(fs), fs.size
    else:
        fsp, nfs = c_void_p(), 0
    mono = None
    if monotone is not None:
        mono = np.zeros(cp.size - 1, dtype=np.int32)
        mono[:len(monotone)] = np.asarray(monotone, dtype=np.int32)
    lib().orc_eval_node_mono(_p(h), GPair64(int(parent_sum[0]), int(parent_sum[1])), _p(cp), _p(cv),
                             c_uint32(cp.size - 1), fsp, c_int32(nfs), byref(param), sc, byref(out), c_void_p(),
                             _p(mono) if mono is not None else c_void_p(), c_float(lower), c_float(upper))
    return out


def eval_node_cat(hist: np.ndarray, parent_sum, cut_ptrs, cut_values, param: Param, scale, feature_is_cat,
                  max_cat_to_onehot: int, max_cat_threshold: int = 64, feature_set=None):
    """Best split over numerical and categorical features (evaluate_splits.cu:148-348).  Returns
    (candidate, is_cat, thresh, set of categories in the split's bit field)."""
    h = np.ascontiguousarray(hist, dtype=np.int64)
    cp = np.ascontiguousarray(cut_ptrs, dtype=np.uint32)
    cv = np.ascontiguousarray(cut_values, dtype=np.float32)
    sc = (c_double * 4)(*[float(v) for v in scale])
    flags = np.ascontiguousarray(feature_is_cat, dtype=np.uint8)
    out = Candidate()
    if feature_set is not None:
        fs = np.ascontiguousarray(feature_set, dtype=np.int32)
        fsp, nfs = _p(fs), fs.size
    else:
        fsp, nfs = c_void_p(), 0
    n_words = int(cv.max()) // 32 + 1 if cv.size else 1
    bits = np.zeros(n_words, dtype=np.uint32)
    is_cat, thresh = c_int32(0), c_int32(-1)
    lib().orc_eval_node_cat(_p(h), GPair64(int(parent_sum[0]), int(parent_sum[1])), _p(cp), _p(cv),
                            c_uint32(cp.size - 1), fsp, c_int32(nfs), byref(param), sc, _p(flags),
                            c_int32(max_cat_to_onehot), c_int32(max_cat_threshold), byref(out), byref(is_cat),
                            byref(thresh), _p(bits), c_uint32(n_words))
    cats = {32 * w + (31 - b) for w in range(n_words) for b in range(32) if (int(bits[w]) >> b) & 1}
    return out, bool(is_cat.value), int(thresh.value), cats


def partition(gbins, cut_ptrs, cut_values, ridx, findex, split_cond, default_left, is_dense_layout=True):
    gb = np.ascontiguousarray(gbins, dtype=np.int32)
    cp = np.ascontiguousarray(cut_ptrs, dtype=np.uint32)
    cv = np.ascontiguousarray(cut_values, dtype=np.float32)
    r = np.ascontiguousarray(ridx, dtype=np.uint32)
    out = np.empty_like(r)
    nl = lib().orc_partition(_p(gb), c_uint32(gb.shape[1]), c_int(1 if is_dense_layout else 0), _p(cp), _p(cv),
                             _p(r), c_uint64(r.size), c_int32(findex), c_float(split_cond),
                             c_int(1 if default_left else 0), _p(out))
    return out, int(nl)


def thrust_shuffle(seed: int, n: int) -> np.ndarray:
    """Permutation thrust::shuffle applies to n elements with default_random_engine(seed): out[i] = input index."""
    perm = np.zeros(max(n, 1), dtype=np.uint32)
    lib().orc_thrust_shuffle(c_uint32(seed & 0xFFFFFFFF), c_uint64(n), _p(perm))
    return perm[:n]


def col_sample(features, colsample: float, seed: int) -> np.ndarray:
    """ColumnSampler::ColSample on CUDA without feature weights: shuffle with the seed, keep int(colsample * n)
    (at least 1), sort ascending."""
    f = np.ascontiguousarray(features, dtype=np.int32)
    out = np.zeros(max(f.size, 1), dtype=np.int32)
    lib().orc_col_sample.restype = c_int32
    k = lib().orc_col_sample(_p(f), c_int32(f.size), c_float(colsample), c_uint32(seed & 0xFFFFFFFF), _p(out))
    return out[:k].copy()


def gradient_based_sample(gq: np.ndarray, scale, subsample: float, seed: int):
    """GradientBasedSampling::Sample restated (pinned on the reference's own gpu_hist trees, see orc_gradient_based_sample).  gq: int64[n, 2]
    quantised pairs.  Returns (sampled pairs, mu, threshold index, regularised absolute gradients)."""
    out = np.ascontiguousarray(gq, dtype=np.int64).copy()
    n = out.shape[0]
    sc = (c_double * 4)(*[float(v) for v in scale])
    tidx = c_uint64(0)
    rag = np.zeros(max(n, 1), dtype=np.float32)
    lib().orc_gradient_based_sample.restype = c_float
    mu = lib().orc_gradient_based_sample(_p(out), c_uint64(n), sc, c_float(subsample), c_uint32(seed & 0xFFFFFFFF),
                                         byref(tidx), _p(rag))
    return out, float(mu), int(tidx.value), rag[:n]


def random_weight(seed: int, i: int) -> float:
    """RandomWeight(seed)(i) of the reference's row sampler (sampler.cu:36-47)."""
    lib().orc_random_weight.restype = c_float
    return float(lib().orc_random_weight(c_uint32(seed & 0xFFFFFFFF), c_uint64(i)))


class _IC(Structure):
    _fields_ = [("n_features", c_int32), ("n_sets", c_int32), ("set_ptr", c_void_p), ("set_feat", c_void_p),
                ("shared_result_buffer", c_int32)]


class InteractionConstraints:
    """FeatureInteractionConstraintDevice (src/tree/constraints.cu) restated: per-node allowed-feature bit fields,
    Split and Query.  `sets` is the parsed `interaction_constraints` parameter, e.g. [[1, 2], [3, 4, 5]]."""

    def __init__(self, sets, n_features: int, max_nodes: int = 64, shared_result_buffer: bool = False):
        self.n_features, self.max_nodes = int(n_features), int(max_nodes)
        self._ptr = np.zeros(len(sets) + 1, dtype=np.int32)
        self._ptr[1:] = np.cumsum([len(s) for s in sets])
        self._feat = np.asarray([f for s in sets for f in s], dtype=np.int32)
        if self._feat.size and (self._feat.min() < 0 or self._feat.max() >= n_features):
            raise ValueError("interaction constraint names a feature outside the matrix")
        self.c = _IC(self.n_features, len(sets), self._ptr.ctypes.data, self._feat.ctypes.data,
                     1 if shared_result_buffer else 0)
        self.words = (self.n_features + 63) // 64
        self.bits = np.zeros((self.max_nodes, max(self.words, 1)), dtype=np.uint64)

    def reset(self):
        self.bits[:] = 0

    def split(self, node: int, fid: int, left: int, right: int):
        assert max(node, left, right) < self.max_nodes and node != left and node != right
        lib().orc_ic_split(byref(self.c), _p(self.bits), c_int32(node), c_int32(fid), c_int32(left), c_int32(right))

    def query_node(self, node: int):
        out = np.zeros(max(self.n_features, 1), dtype=np.int32)
        n = lib().orc_ic_query_node(byref(self.c), _p(self.bits), c_int32(node), _p(out))
        return out[:n].tolist()

    def query(self, feature_list, node: int):
        fl = np.ascontiguousarray(feature_list, dtype=np.int32)
        out = np.zeros(max(self.n_features, fl.size, 1), dtype=np.int32)
        n = lib().orc_ic_query(byref(self.c), _p(self.bits), _p(fl), c_int32(fl.size), c_int32(node), _p(out))
        return out[:n].tolist()


class _Cat(Structure):
    _fields_ = [("feature_is_cat", c_void_p), ("max_cat_to_onehot", c_int32), ("max_cat_threshold", c_int32),
                ("n_cat_words", c_uint32), ("node_is_cat", c_void_p), ("node_cat_bits", c_void_p),
                ("node_sets", c_void_p), ("n_node_sets", c_int32), ("n_per_node", c_int32), ("eval_count", c_void_p),
                ("gradient_based", c_int32), ("seeds", c_void_p), ("n_seeds", c_int32), ("tree_set", c_void_p),
                ("n_tree_set", c_int32), ("colsample_bylevel", c_float), ("colsample_bynode", c_float),
                ("draws_used", c_void_p)]


def update_tree(gbins, cut_ptrs, cut_values, gpair, param: Param, scale_override=None, max_nodes=None,
                is_dense_layout=True, want_hist=False, feature_set=None, subsample=1.0, sample_seed=0,
                monotone=None, level_feature_sets=None, interaction=None, interaction_shared_buffer=True,
                feature_is_cat=None, max_cat_to_onehot=4, max_cat_threshold=64, node_feature_sets=None,
                sampling_method="uniform", column_seeds=None, tree_set=None, colsample_bylevel=1.0, colsample_bynode=1.0,
                node_batch=None):
    """Whole tree.  Returns dict(nodes=[Node...], position, scale, hist (optional), cands).
    node_batch: the Driver's max_node_batch_size for this call (default: the GPU updater's 1024).
    feature_set: sorted feature ids of the tree's colsample_bytree draw; subsample/sample_seed: uniform row sampling;
    monotone: per-feature sign (+1 increasing, -1 decreasing, 0 free), padded with 0 to n_features;
    feature_is_cat: flags per feature (categorical features: bins are categories, cut values their codes), with
    TrainParam's max_cat_to_onehot / max_cat_threshold; the result then carries is_cat[n_nodes] and cat_sets (the
    categories that go RIGHT, per categorical split);
    column_seeds (+ tree_set, colsample_bylevel, colsample_bynode): the column sampler driven by the engine's outputs --
    level sets drawn lazily per depth, a node set per evaluated node, in evaluation order (result: draws_used);
    node_feature_sets: colsample_bynode, int32[n_draws, n_per_node]: the k-th node evaluated takes row k (the result
    then carries eval_count, the number of draws used);
    interaction: parsed interaction_constraints ([[f, ...], ...]); interaction_shared_buffer (default = what the
    reference's trees follow, tests/golden/ref_gpu_interaction_trees.json): see orc_update_tree_ic."""
    gb = np.ascontiguousarray(gbins, dtype=np.int32)
    n, stride = gb.shape
    cp = np.ascontiguousarray(cut_ptrs, dtype=np.uint32)
    cv = np.ascontiguousarray(cut_values, dtype=np.float32)
    g = np.ascontiguousarray(gpair, dtype=np.float32)
    nf = cp.size - 1
    if max_nodes is None:
        by_d = (1 << (param.max_depth + 1)) - 1 if param.max_depth > 0 else 1 << 30
        by_l = 2 * param.max_leaves - 1 if param.max_leaves > 0 else 1 << 30
        max_nodes = min(by_d, by_l)
    nodes = (Node * max_nodes)()
    cands = (Candidate * max_nodes)()
    pos = np.empty(n, dtype=np.int32)
    scale_out = (c_double * 4)()
    so = (c_double * 4)(*[float(v) for v in scale_override]) if scale_override is not None else None
    hist = np.zeros((max_nodes, cv.size, 2), dtype=np.int64) if want_hist else None
    fs = np.ascontiguousarray(feature_set, dtype=np.int32) if feature_set is not None else None
    n_levels, n_fs = 0, (0 if fs is None else fs.size)
    if level_feature_sets is not None:   # colsample_bylevel: [n_levels][n_per_level], row = node depth
        fs = np.ascontiguousarray(level_feature_sets, dtype=np.int32)
        n_levels, n_fs = fs.shape
    mono = None
    if monotone is not None:
        mono = np.zeros(nf, dtype=np.int32)
        mono[:len(monotone)] = np.asarray(monotone, dtype=np.int32)
    ic = InteractionConstraints(interaction, nf, 1, interaction_shared_buffer) if interaction is not None else None
    cat = None
    has_cat = feature_is_cat is not None and any(feature_is_cat)
    eval_count = np.zeros(1, dtype=np.int32)
    gradient_based = sampling_method == "gradient_based" and subsample < 1.0
    draws_used = np.zeros(1, dtype=np.int32)
    if has_cat or node_feature_sets is not None or gradient_based or column_seeds is not None:
        cat = _Cat(None, int(max_cat_to_onehot), int(max_cat_threshold), 1, None, None, None, 0, 0, eval_count.ctypes.data,
                   1 if gradient_based else 0, None, 0, None, 0, 1.0, 1.0, draws_used.ctypes.data)
    if column_seeds is not None:
        sd = np.ascontiguousarray([int(v) & 0xFFFFFFFF for v in column_seeds], dtype=np.uint32)
        ts_ = np.ascontiguousarray(tree_set if tree_set is not None else np.arange(nf), dtype=np.int32)
        cat.seeds, cat.n_seeds, cat.tree_set, cat.n_tree_set = sd.ctypes.data, sd.size, ts_.ctypes.data, ts_.size
        cat.colsample_bylevel, cat.colsample_bynode = float(colsample_bylevel), float(colsample_bynode)
    if has_cat:
        flags = np.zeros(nf, dtype=np.uint8)
        flags[:len(feature_is_cat)] = [1 if v else 0 for v in feature_is_cat]
        n_words = int(cv.max()) // 32 + 1 if cv.size else 1
        node_is_cat = np.zeros(max_nodes, dtype=np.int32)
        node_bits = np.zeros((max_nodes, n_words), dtype=np.uint32)
        cat.feature_is_cat, cat.n_cat_words = flags.ctypes.data, n_words
        cat.node_is_cat, cat.node_cat_bits = node_is_cat.ctypes.data, node_bits.ctypes.data
    if node_feature_sets is not None:
        ns = np.ascontiguousarray(node_feature_sets, dtype=np.int32)
        cat.node_sets, cat.n_node_sets, cat.n_per_node = ns.ctypes.data, ns.shape[0], ns.shape[1]
    lib().orc_update_tree_full.restype = c_int32
    lib().orc_set_node_batch(c_int32(int(node_batch) if node_batch else 0))   # 0 restores the default (1024)
    nn = lib().orc_update_tree_full(_p(gb), c_uint64(n), c_uint32(stride), c_int(1 if is_dense_layout else 0),
                                  c_uint32(nf), _p(cp), _p(cv), _p(g), byref(param), so, nodes, c_int32(max_nodes),
                                  _p(pos), _p(hist) if want_hist else c_void_p(), cands, scale_out,
                                  _p(fs) if fs is not None else c_void_p(), c_int32(n_fs),
                                  c_float(subsample), c_uint32(int(sample_seed) & 0xFFFFFFFF),
                                  _p(mono) if mono is not None else c_void_p(), c_int32(n_levels),
                                  byref(ic.c) if ic is not None else c_void_p(),
                                  byref(cat) if cat is not None else c_void_p())
    lib().orc_set_node_batch(c_int32(0))
    res = dict(nodes=list(nodes[:nn]), cands=list(cands[:nn]), position=pos,
               scale=np.array(scale_out[:], dtype=np.float64), hist=hist[:nn] if want_hist else None)
    if node_feature_sets is not None:
        res["eval_count"] = int(eval_count[0])
    if column_seeds is not None:
        res["draws_used"] = int(draws_used[0])
    if has_cat:
        res["is_cat"] = node_is_cat[:nn].copy()
        res["cat_sets"] = [{32 * w + (31 - b) for w in range(n_words) for b in range(32) if (int(node_bits[i, w]) >> b) & 1}
                           if node_is_cat[i] else set() for i in range(nn)]
    return res


def predict_leaf(gbins, cut_ptrs, cut_values, nodes, is_cat=None, cat_sets=None, is_dense_layout=True) -> np.ndarray:
    """Leaf every row reaches in ONE tree (nodes: list of Node, or anything with the same fields).  cat_sets: per
    node, the set of categories that go right (for nodes flagged in is_cat)."""
    gb = np.ascontiguousarray(gbins, dtype=np.int32)
    cp = np.ascontiguousarray(cut_ptrs, dtype=np.uint32)
    cv = np.ascontiguousarray(cut_values, dtype=np.float32)
    arr = (Node * len(nodes))(*nodes)
    out = np.empty(gb.shape[0], dtype=np.int32)
    flags = bits = None
    n_words = 1
    if is_cat is not None and any(is_cat):
        n_words = int(cv.max()) // 32 + 1
        flags = np.ascontiguousarray(is_cat, dtype=np.int32)
        bits = np.zeros((len(nodes), n_words), dtype=np.uint32)
        for i, cs in enumerate(cat_sets):
            for c in cs:
                bits[i, c // 32] |= np.uint32(1 << (31 - c % 32))
    lib().orc_predict_leaf(_p(gb), c_uint64(gb.shape[0]), c_uint32(gb.shape[1]), c_int(1 if is_dense_layout else 0),
                           _p(cp), _p(cv), arr, _p(flags) if flags is not None else c_void_p(),
                           _p(bits) if bits is not None else c_void_p(), c_uint32(n_words), _p(out))
    return out


def predict_margin(gbins, cut_ptrs, cut_values, trees, base=None, is_dense_layout=True) -> np.ndarray:
    """GPUPredictor::PredictBatch on binned data (gpu_predictor.cu:221-228): per row, float sum = 0; sum += leaf of
    every tree in order; out += sum.  trees: list of dict(nodes=[Node...], is_cat=..., cat_sets=...) as returned by
    update_tree, or plain node lists."""
    n = np.asarray(gbins).shape[0]
    total = np.zeros(n, dtype=np.float32)
    for t in trees:
        nodes = t["nodes"] if isinstance(t, dict) else t
        leaf = predict_leaf(gbins, cut_ptrs, cut_values, nodes, t.get("is_cat") if isinstance(t, dict) else None,
                            t.get("cat_sets") if isinstance(t, dict) else None, is_dense_layout)
        values = np.array([nd.leaf_value for nd in nodes], dtype=np.float32)
        total = (total + values[leaf]).astype(np.float32)
    out = np.zeros(n, dtype=np.float32) if base is None else np.asarray(base, dtype=np.float32).copy()
    return (out + total).astype(np.float32)


def predict_margin_groups(gbins, cut_ptrs, cut_values, trees, tree_groups, n_groups, base=None,
                          is_dense_layout=True) -> np.ndarray:
    """GPUPredictor's kernel with more than one output group (gpu_predictor.cu:229-239; the CPU predictor does the same,
    cpu_predictor.cc PredictByAllTrees): every leaf is added STRAIGHT onto out[row, group(tree)], tree after tree --
    no sum from 0 first.  Returns float32 [n_rows, n_groups]."""
    n = np.asarray(gbins).shape[0]
    out = (np.zeros((n, n_groups), dtype=np.float32) if base is None
           else np.asarray(base, dtype=np.float32).reshape(n, n_groups).copy())
    for t, g in zip(trees, tree_groups):
        nodes = t["nodes"] if isinstance(t, dict) else t
        leaf = predict_leaf(gbins, cut_ptrs, cut_values, nodes, t.get("is_cat") if isinstance(t, dict) else None,
                            t.get("cat_sets") if isinstance(t, dict) else None, is_dense_layout)
        values = np.array([nd.leaf_value for nd in nodes], dtype=np.float32)
        out[:, g] = (out[:, g] + values[leaf]).astype(np.float32)
    return out


def update_prediction_cache(nodes, position: np.ndarray, preds: np.ndarray) -> np.ndarray:
    arr = (Node * len(nodes))(*nodes)
    p = np.ascontiguousarray(preds, dtype=np.float32).copy()
    pos = np.ascontiguousarray(position, dtype=np.int32)
    lib().orc_update_prediction_cache(arr, _p(pos), c_uint64(pos.size), _p(p))
    return p
