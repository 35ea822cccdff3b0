"""ctypes binding of the C ABI in include/b200hist.h.

The library is mandatory: importing this module on a machine where libb200hist.so is missing, or
calling any compute entry point without a CUDA device, raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64,
                    c_uint8, c_uint32, c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200hist.so")


class B200HistError(RuntimeError):
    pass


class TrainParam(Structure):
    """b200hist_train_param; defaults are the reference's TrainParam defaults (src/tree/param.h)."""
    _fields_ = [("max_depth", c_int32), ("max_leaves", c_int32), ("grow_policy", c_int32),
                ("learning_rate", c_float), ("reg_lambda", c_float), ("reg_alpha", c_float),
                ("min_child_weight", c_float), ("max_delta_step", c_float), ("min_split_loss", c_float),
                ("reserved", c_int32 * 7)]

    @classmethod
    def make(cls, **kw):
        p = cls(max_depth=6, max_leaves=0, grow_policy=0, learning_rate=0.3, reg_lambda=1.0, reg_alpha=0.0,
                min_child_weight=1.0, max_delta_step=0.0, min_split_loss=0.0)
        alias = {"eta": "learning_rate", "lambda": "reg_lambda", "alpha": "reg_alpha", "gamma": "min_split_loss"}
        for k, v in kw.items():
            k = alias.get(k, k)
            if k == "grow_policy" and isinstance(v, str):
                v = {"depthwise": 0, "lossguide": 1}[v]
            if not hasattr(p, k):
                raise B200HistError("unknown TrainParam field %r" % k)
            setattr(p, k, v)
        return p


class GPair64(Structure):
    _fields_ = [("grad", c_int64), ("hess", c_int64)]


class Quantiser(Structure):
    _fields_ = [("to_fixed_grad", c_double), ("to_fixed_hess", c_double),
                ("to_float_grad", c_double), ("to_float_hess", c_double)]


class Candidate(Structure):
    _fields_ = [("loss_chg", c_float), ("dir", c_int32), ("findex", c_int32), ("fvalue", c_float),
                ("left_sum", GPair64), ("right_sum", GPair64),
                ("base_weight", c_float), ("left_weight", c_float), ("right_weight", c_float),
                ("split_gbin", c_int32)]


class TreeNode(Structure):
    _fields_ = [("parent", c_int32), ("left", c_int32), ("right", c_int32), ("findex", c_int32),
                ("fvalue", c_float), ("default_left", c_int32), ("is_leaf", c_int32), ("leaf_value", c_float),
                ("loss_chg", c_float), ("sum_hess", c_float), ("base_weight", c_float), ("depth", c_int32)]


_SIGNATURES = {
    "b200hist_last_error": (c_char_p, []),
    "b200hist_version": (c_int, []),
    "b200hist_launch_count": (c_uint64, []),
    "b200hist_ellpack_create": (c_int, [c_int, c_uint64, c_uint32, POINTER(c_uint32), POINTER(c_float), c_int,
                                        POINTER(c_void_p)]),
    "b200hist_sketch_dense": (c_int, [c_int, c_void_p, c_uint64, c_uint64, c_uint32, c_int32, POINTER(c_uint32),
                                      POINTER(c_float), c_uint64, c_void_p]),
    "b200hist_updater_last_comm_ms": (c_int, [c_void_p, POINTER(c_float)]),
    "b200hist_updater_set_tree_sampling": (c_int, [c_void_p, POINTER(c_int32), c_uint32, c_float, c_uint32]),
    "b200hist_updater_set_level_feature_sets": (c_int, [c_void_p, POINTER(c_int32), c_uint32, c_uint32]),
    "b200hist_updater_last_max_eval_depth": (c_int, [c_void_p, POINTER(c_int32)]),
    "b200hist_updater_set_monotone_constraints": (c_int, [c_void_p, POINTER(c_int32), c_uint32]),
    "b200hist_updater_set_interaction_constraints": (c_int, [c_void_p, POINTER(c_int32), POINTER(c_int32), c_uint32]),
    "b200hist_updater_set_sampling_method": (c_int, [c_void_p, c_int]),
    "b200hist_updater_set_column_sampler": (c_int, [c_void_p, POINTER(c_int32), c_uint32, c_float, c_float,
                                                    POINTER(c_uint32), c_uint32]),
    "b200hist_updater_last_draw_count": (c_int, [c_void_p, POINTER(c_int32)]),
    "b200hist_col_sample_stream": (c_int, [c_int, POINTER(c_int32), c_uint32, c_float, POINTER(c_uint32), c_uint32,
                                           POINTER(c_int32), POINTER(c_uint32), c_void_p]),
    "b200hist_updater_set_node_feature_sets": (c_int, [c_void_p, POINTER(c_int32), c_uint32, c_uint32]),
    "b200hist_updater_last_eval_count": (c_int, [c_void_p, POINTER(c_int32)]),
    "b200hist_updater_set_categorical": (c_int, [c_void_p, c_void_p, c_uint32, c_int32, c_int32]),
    "b200hist_updater_cat_words": (c_int, [c_void_p, POINTER(c_uint32)]),
    "b200hist_updater_last_tree_categories": (c_int, [c_void_p, POINTER(c_int32), POINTER(c_uint32)]),
    "b200hist_updater_clear_peers": (c_int, [c_void_p]),
    "b200hist_updater_set_row_split": (c_int, [c_void_p, c_int]),
    "b200hist_smem_atomic_peak": (c_int, [c_int, c_int, POINTER(c_double), c_void_p]),
    "b200hist_sketch_create": (c_int, [c_int, c_uint32, c_int32, POINTER(c_void_p)]),
    "b200hist_sketch_push_dense": (c_int, [c_void_p, c_void_p, c_uint64, c_uint64, c_void_p]),
    "b200hist_sketch_set_feature_types": (c_int, [c_void_p, c_void_p, c_uint32]),
    "b200hist_sketch_make_cuts": (c_int, [c_void_p, POINTER(c_uint32), POINTER(c_float), c_uint64, POINTER(c_uint32),
                                          c_void_p]),
    "b200hist_sketch_destroy": (c_int, [c_void_p]),
    "b200hist_ellpack_create_sparse": (c_int, [c_int, c_uint64, c_uint32, POINTER(c_uint32), POINTER(c_float), c_uint32,
                                               POINTER(c_void_p)]),
    "b200hist_ellpack_push_csr": (c_int, [c_void_p, c_uint64, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]),
    "b200hist_ellpack_destroy": (c_int, [c_void_p]),
    "b200hist_ellpack_create_paged": (c_int, [c_int, c_uint64, c_uint32, POINTER(c_uint32), POINTER(c_float), c_int,
                                              c_uint64, POINTER(c_void_p)]),
    "b200hist_ellpack_paging_info": (c_int, [c_void_p, POINTER(c_uint32), POINTER(c_uint64), POINTER(c_uint64)]),
    "b200hist_ellpack_set_feature_types": (c_int, [c_void_p, c_void_p, c_uint32]),
    "b200hist_ellpack_push_dense": (c_int, [c_void_p, c_uint64, c_void_p, c_uint64, c_uint64, c_void_p]),
    "b200hist_ellpack_import_reference": (c_int, [c_void_p, c_void_p, c_uint32, c_uint32, c_int, c_void_p]),
    "b200hist_ellpack_import_reference_rows": (c_int, [c_void_p, c_uint64, c_uint64, c_void_p, c_uint32, c_uint32, c_int,
                                                       c_void_p]),
    "b200hist_ellpack_info": (c_int, [c_void_p, POINTER(c_uint64), POINTER(c_uint32), POINTER(c_uint32),
                                      POINTER(c_uint32), POINTER(c_uint32), POINTER(c_uint32), POINTER(c_void_p)]),
    "b200hist_ellpack_copy_bins_to_host": (c_int, [c_void_p, c_void_p]),
    "b200hist_updater_create": (c_int, [c_void_p, POINTER(TrainParam), POINTER(c_void_p)]),
    "b200hist_updater_destroy": (c_int, [c_void_p]),
    "b200hist_updater_max_nodes": (c_int, [c_void_p, POINTER(c_int32)]),
    "b200hist_updater_update_tree": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(TreeNode), POINTER(c_int32),
                                             c_void_p]),
    "b200hist_updater_update_tree_host": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(TreeNode),
                                                  POINTER(c_int32), c_void_p]),
    "b200hist_updater_update_tree_async": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200hist_updater_fetch_tree": (c_int, [c_void_p, POINTER(TreeNode), POINTER(c_int32)]),
    "b200hist_updater_update_prediction_cache": (c_int, [c_void_p, c_void_p, c_void_p]),
    "b200hist_updater_update_prediction_cache_ex": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_void_p, c_int32,
                                                            c_void_p]),
    "b200hist_comm_unique_id": (c_int, [c_void_p]),
    "b200hist_comm_create": (c_int, [c_int, c_int, c_int, c_void_p, POINTER(c_void_p)]),
    "b200hist_comm_destroy": (c_int, [c_void_p]),
    "b200hist_shard_groups": (c_int, [c_uint32, c_int, c_int, POINTER(c_uint32), POINTER(c_uint32)]),
    "b200hist_updater_set_comm": (c_int, [c_void_p, c_void_p]),
    "b200hist_updater_num_feature_groups": (c_int, [c_void_p, POINTER(c_uint32)]),
    "b200hist_updater_ipc_handle": (c_int, [c_void_p, c_void_p]),
    "b200hist_updater_set_peers": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "b200hist_updater_last_quantiser": (c_int, [c_void_p, POINTER(Quantiser)]),
    "b200hist_updater_node_histogram": (c_int, [c_void_p, c_int32, c_void_p]),
    "b200hist_updater_node_candidate": (c_int, [c_void_p, c_int32, POINTER(Candidate)]),
    "b200hist_updater_set_profiling": (c_int, [c_void_p, c_int]),
    "b200hist_updater_last_hist_steps": (c_int, [c_void_p, c_void_p, c_int32, POINTER(c_int32)]),
    "b200hist_updater_last_profile": (c_int, [c_void_p, POINTER(c_float), POINTER(c_uint64)]),
    "b200hist_quantise": (c_int, [c_int, c_void_p, c_uint64, POINTER(Quantiser), c_void_p, POINTER(GPair64),
                                  c_void_p]),
    "b200hist_build_hist": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_uint64), POINTER(c_uint64), c_int32,
                                    c_void_p, c_void_p]),
    "b200hist_subtract": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]),
    "b200hist_evaluate_splits": (c_int, [c_void_p, c_void_p, POINTER(GPair64), c_int32, POINTER(c_int32), c_int32,
                                         POINTER(TrainParam), POINTER(Quantiser), POINTER(Candidate), c_void_p]),
    "b200hist_evaluate_splits_cat": (c_int, [c_void_p, c_void_p, POINTER(GPair64), c_int32, POINTER(c_int32), c_int32,
                                             POINTER(TrainParam), POINTER(Quantiser), c_void_p, c_int32, c_int32,
                                             POINTER(Candidate), POINTER(c_int32), POINTER(c_uint32), c_uint32,
                                             c_void_p]),
    "b200hist_forest_create": (c_int, [c_void_p, POINTER(c_void_p)]),
    "b200hist_forest_destroy": (c_int, [c_void_p]),
    "b200hist_forest_num_trees": (c_int, [c_void_p, POINTER(c_uint32)]),
    "b200hist_forest_add_tree": (c_int, [c_void_p, c_void_p, c_int32, POINTER(c_int32), POINTER(c_uint32), c_uint32]),
    "b200hist_forest_predict_margin": (c_int, [c_void_p, c_void_p, c_uint32, c_uint32, c_uint64, c_uint64, c_void_p,
                                               c_void_p]),
    "b200hist_forest_predict_leaf": (c_int, [c_void_p, c_void_p, c_uint32, c_uint32, c_uint64, c_uint64, c_void_p,
                                             c_void_p]),
    "b200hist_forest_predict_leaf_f32": (c_int, [c_void_p, c_void_p, c_uint32, c_uint32, c_uint64, c_uint64, c_void_p,
                                                 c_void_p]),
    "b200hist_forest_predict_reference_page": (c_int, [c_void_p, c_void_p, c_uint64, c_uint32, c_uint32, c_int, c_int,
                                                       c_int, c_uint32, c_uint32, c_uint32, c_void_p, c_void_p]),
    "b200hist_forest_set_tree_groups": (c_int, [c_void_p, POINTER(c_uint32), c_uint32]),
    "b200hist_forest_predict_margin_groups": (c_int, [c_void_p, c_void_p, c_uint32, c_uint32, c_uint32, c_uint64,
                                                      c_uint64, c_void_p, c_void_p]),
    "b200hist_partition": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_uint64), POINTER(c_uint64),
                                   POINTER(c_int32), POINTER(c_float), POINTER(c_int32), c_int32,
                                   POINTER(c_uint64), c_void_p]),
}

# every symbol include/b200hist.h declares
EXPORTED_SYMBOLS = tuple(_SIGNATURES.keys())

_lib = None


def load() -> ctypes.CDLL:
    """Load libb200hist.so; fails loudly if it has not been built (python -m xgboost_b200._build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200HistError(
                "libb200hist.so is missing at %s -- build it with `python xgboost_b200/_build.py` "
                "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise B200HistError(load().b200hist_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(load().b200hist_launch_count())
