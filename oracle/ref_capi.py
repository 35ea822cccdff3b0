"""Minimal ctypes driver of the UNMODIFIED reference library (libxgboost.so built from
/root/reference by oracle/Makefile.ref or oracle/Makefile.ref_cuda) through its public C API
(/root/reference/include/xgboost/c_api.h).

TEST / BENCH INFRASTRUCTURE ONLY.  Used to (1) generate golden fixtures (tools/make_golden*.py),
(2) time the reference CPU `hist` path (bench.py --impl reference, cpu_baseline) and (3) time the
reference CUDA gpu_hist next to ours.  The product never imports this.
"""
from __future__ import annotations

import ctypes
import json
import os
from ctypes import POINTER, byref, c_char_p, c_float, c_int, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CPU_LIB = os.path.join(_HERE, "_ref", "libxgboost.so")
CUDA_LIB = os.path.join(_HERE, "_ref_cuda", "libxgboost.so")

_libs = {}


def available(kind: str = "cpu") -> bool:
    return os.path.exists(CPU_LIB if kind == "cpu" else CUDA_LIB)


def load(kind: str = "cpu") -> ctypes.CDLL:
    if kind not in _libs:
        path = CPU_LIB if kind == "cpu" else CUDA_LIB
        if not os.path.exists(path):
            raise RuntimeError("reference library %s not built (make -C oracle -f Makefile.ref%s)" %
                               (path, "" if kind == "cpu" else "_cuda"))
        lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        lib.XGBGetLastError.restype = c_char_p
        _libs[kind] = lib
    return _libs[kind]


class RefError(RuntimeError):
    pass


def _check(lib, rc):
    if rc != 0:
        raise RefError(lib.XGBGetLastError().decode())


def _array_interface(a: np.ndarray) -> bytes:
    d = a.__array_interface__
    return json.dumps({"data": [d["data"][0], False], "shape": list(a.shape), "typestr": d["typestr"],
                       "version": 3}).encode()


def _cuda_array_interface(t) -> bytes:
    d = t.__cuda_array_interface__
    return json.dumps({"data": [d["data"][0], False], "shape": list(d["shape"]), "typestr": d["typestr"],
                       "version": 3, "stream": 1}).encode()


class DMatrix:
    def __init__(self, lib, handle, keep=None):
        self.lib, self.handle, self._keep = lib, handle, keep

    @classmethod
    def from_dense(cls, x: np.ndarray, label=None, kind="cpu", nthread=0):
        lib = load(kind)
        x = np.ascontiguousarray(x, dtype=np.float32)
        h = c_void_p()
        cfg = json.dumps({"missing": float("nan"), "nthread": int(nthread)}).encode()
        _check(lib, lib.XGDMatrixCreateFromDense(_array_interface(x), cfg, byref(h)))
        dm = cls(lib, h, keep=x)
        if label is not None:
            dm.set_label(label)
        return dm

    @classmethod
    def from_csr(cls, indptr, indices, values, n_cols, label=None, kind="cpu", nthread=0):
        """XGDMatrixCreateFromCSR (c_api.h): host CSR arrays."""
        lib = load(kind)
        ip = np.ascontiguousarray(indptr, dtype=np.uint64)
        ix = np.ascontiguousarray(indices, dtype=np.uint32)
        va = np.ascontiguousarray(values, dtype=np.float32)
        h = c_void_p()
        cfg = json.dumps({"missing": float("nan"), "nthread": int(nthread)}).encode()
        _check(lib, lib.XGDMatrixCreateFromCSR(_array_interface(ip), _array_interface(ix), _array_interface(va),
                                               c_uint64(n_cols), cfg, byref(h)))
        dm = cls(lib, h, keep=(ip, ix, va))
        if label is not None:
            dm.set_label(label)
        return dm

    @classmethod
    def from_cuda(cls, x_cuda, label=None, kind="cuda"):
        """x_cuda: torch CUDA float32 matrix (in-core SimpleDMatrix on the device)."""
        lib = load(kind)
        h = c_void_p()
        cfg = json.dumps({"missing": float("nan"), "nthread": 0}).encode()
        _check(lib, lib.XGDMatrixCreateFromCudaArrayInterface(_cuda_array_interface(x_cuda), cfg, byref(h)))
        dm = cls(lib, h, keep=x_cuda)
        if label is not None:
            dm.set_label(label)
        return dm

    @classmethod
    def quantile_from_batches(cls, batches, max_bin=256, kind="cpu", nthread=0):
        """QuantileDMatrix from a list of batches (numpy arrays for the CPU library, torch CUDA tensors for
        the CUDA one) via XGQuantileDMatrixCreateFromCallback (c_api.h:538)."""
        lib = load(kind)
        proxy = c_void_p()
        _check(lib, lib.XGProxyDMatrixCreate(byref(proxy)))
        state = {"i": 0, "cur": None}
        if callable(batches):  # (n_batches, get_batch) form: batches are regenerated on every pass
            raise TypeError("pass a sequence; use LazyBatches for regenerated batches")
        first = batches[0]
        is_cuda = hasattr(first, "__cuda_array_interface__") and not isinstance(first, np.ndarray)

        @ctypes.CFUNCTYPE(None, c_void_p)
        def reset(_):
            state["i"] = 0

        @ctypes.CFUNCTYPE(c_int, c_void_p)
        def nxt(_):
            i = state["i"]
            if i >= len(batches):
                return 0
            b = batches[i]
            state["cur"] = b  # keep the batch alive while the library reads it
            if is_cuda:
                rc = lib.XGProxyDMatrixSetDataCudaArrayInterface(proxy, _cuda_array_interface(b))
            else:
                rc = lib.XGProxyDMatrixSetDataDense(proxy, _array_interface(b))
            if rc != 0:
                return -1
            state["i"] = i + 1
            return 1

        cfg = json.dumps({"missing": float("nan"), "nthread": int(nthread), "max_bin": int(max_bin)}).encode()
        h = c_void_p()
        _check(lib, lib.XGQuantileDMatrixCreateFromCallback(None, proxy, None, reset, nxt, cfg, byref(h)))
        return cls(lib, h, keep=(batches, reset, nxt, proxy))

    @classmethod
    def extmem_quantile_from_batches(cls, batches, max_bin=256, kind="cuda", on_host=True, nthread=0):
        """ExtMemQuantileDMatrix over a list of batches via XGExtMemQuantileDMatrixCreateFromCallback (c_api.h:579):
        one Ellpack page per batch (min_cache_page_bytes = 0 disables page concatenation), cached on the host."""
        lib = load(kind)
        proxy = c_void_p()
        _check(lib, lib.XGProxyDMatrixCreate(byref(proxy)))
        state = {"i": 0, "cur": None}
        first = batches[0]
        is_cuda = hasattr(first, "__cuda_array_interface__") and not isinstance(first, np.ndarray)

        @ctypes.CFUNCTYPE(None, c_void_p)
        def reset(_):
            state["i"] = 0

        @ctypes.CFUNCTYPE(c_int, c_void_p)
        def nxt(_):
            i = state["i"]
            if i >= len(batches):
                return 0
            b = batches[i]
            state["cur"] = b
            if is_cuda:
                rc = lib.XGProxyDMatrixSetDataCudaArrayInterface(proxy, _cuda_array_interface(b))
            else:
                rc = lib.XGProxyDMatrixSetDataDense(proxy, _array_interface(b))
            if rc != 0:
                return -1
            state["i"] = i + 1
            return 1

        cfg = json.dumps({"missing": float("nan"), "nthread": int(nthread), "max_bin": int(max_bin),
                          "cache_prefix": "b200hist_extmem_cache", "on_host": bool(on_host),
                          "min_cache_page_bytes": 0}).encode()
        h = c_void_p()
        _check(lib, lib.XGExtMemQuantileDMatrixCreateFromCallback(None, proxy, None, reset, nxt, cfg, byref(h)))
        return cls(lib, h, keep=(batches, reset, nxt, proxy))

    def set_label(self, y):
        y = np.ascontiguousarray(y, dtype=np.float32)
        _check(self.lib, self.lib.XGDMatrixSetInfoFromInterface(self.handle, b"label", _array_interface(y)))

    def set_feature_types(self, types):
        """XGDMatrixSetStrFeatureInfo(handle, "feature_type", ...): "q" numerical, "c" categorical, ..."""
        arr = (c_char_p * len(types))(*[t.encode() for t in types])
        _check(self.lib, self.lib.XGDMatrixSetStrFeatureInfo(self.handle, b"feature_type", arr, c_uint64(len(types))))
        self._feature_types = list(types)

    def quantile_cut(self):
        """(cut_ptrs uint32/64, cut_values float32) via XGDMatrixGetQuantileCut (c_api.h:972)."""
        ip, dt = c_char_p(), c_char_p()
        _check(self.lib, self.lib.XGDMatrixGetQuantileCut(self.handle, b"{}", byref(ip), byref(dt)))

        return self._to_numpy(json.loads(ip.value)), self._to_numpy(json.loads(dt.value))

    def histogram_cuts(self):
        """HistogramCuts (cut_ptrs uint32, cut_values float32): XGDMatrixGetQuantileCut prepends each
        numerical feature's lower bound (src/c_api/c_api.cc:967-994); strip it."""
        ip, dv = self.quantile_cut()
        ptrs, vals = [0], []
        types = getattr(self, "_feature_types", None)
        for f in range(len(ip) - 1):
            is_cat = types is not None and types[f] == "c"   # no lower bound is prepended for a categorical feature
            seg = dv[int(ip[f]) + (0 if is_cat else 1):int(ip[f + 1])]
            vals.append(seg)
            ptrs.append(ptrs[-1] + seg.size)
        return np.array(ptrs, dtype=np.uint32), np.concatenate(vals).astype(np.float32)

    @staticmethod
    def _to_numpy(d):
        n = int(np.prod(d["shape"]))
        dtype = np.dtype(d["typestr"])
        ptr = d["data"][0]
        out = np.empty(n, dtype=dtype)
        if "stream" in d:  # __cuda_array_interface__: device pointer
            import torch
            rc = torch.cuda.cudart().cudaMemcpy(out.ctypes.data, ptr, n * dtype.itemsize, 2)
            if int(rc) != 0:
                raise RefError("cudaMemcpy of the quantile cuts failed")
        else:
            ctypes.memmove(out.ctypes.data, ptr, n * dtype.itemsize)
        return out

    def __del__(self):
        try:
            self.lib.XGDMatrixFree(self.handle)
        except Exception:
            pass


class LazyBatches:
    """Sequence whose items are produced on demand (so a 100M-row matrix never exists at once)."""

    def __init__(self, n: int, get):
        self.n, self.get = n, get

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if i < 0 or i >= self.n:
            raise IndexError(i)
        return self.get(i)


class Booster:
    def __init__(self, params: dict, dtrain: DMatrix):
        self.lib = dtrain.lib
        self.h = c_void_p()
        arr = (c_void_p * 1)(dtrain.handle)
        _check(self.lib, self.lib.XGBoosterCreate(arr, c_uint64(1), byref(self.h)))
        for k, v in params.items():
            _check(self.lib, self.lib.XGBoosterSetParam(self.h, str(k).encode(), str(v).encode()))
        self.dtrain = dtrain

    def update(self, it: int):
        _check(self.lib, self.lib.XGBoosterUpdateOneIter(self.h, c_int(it), self.dtrain.handle))

    def boost(self, it: int, grad, hess):
        """XGBoosterTrainOneIter with custom gradients (numpy for host, torch CUDA for device)."""
        if isinstance(grad, np.ndarray):
            g, h = _array_interface(grad), _array_interface(hess)
        else:
            g, h = _cuda_array_interface(grad), _cuda_array_interface(hess)
        _check(self.lib, self.lib.XGBoosterTrainOneIter(self.h, self.dtrain.handle, c_int(it), g, h))

    def dump(self):
        n = c_uint64()
        out = POINTER(c_char_p)()
        _check(self.lib, self.lib.XGBoosterDumpModelEx(self.h, b"", c_int(1), b"json", byref(n), byref(out)))
        return [json.loads(out[i].decode()) for i in range(n.value)]

    def save_json(self) -> dict:
        n = c_uint64()
        p = POINTER(ctypes.c_char)()
        _check(self.lib, self.lib.XGBoosterSaveModelToBuffer(self.h, b'{"format": "json"}', byref(n), byref(p)))
        return json.loads(ctypes.string_at(p, n.value).decode())

    def predict_margin(self, dm: DMatrix, pred_type: int = 1, iteration_begin: int = 0) -> np.ndarray:
        """pred_type: 1 = margin, 6 = leaf ids (c_api.h XGBoosterPredictFromDMatrix); flat float32.
        iteration_begin > 0 also bypasses the Booster's prediction cache (GBTree::PredictBatchImpl resets it)."""
        cfg = json.dumps({"type": int(pred_type), "training": False, "iteration_begin": int(iteration_begin), "iteration_end": 0,
                          "strict_shape": False}).encode()
        shape = POINTER(c_uint64)()
        dim = c_uint64()
        res = POINTER(c_float)()
        _check(self.lib, self.lib.XGBoosterPredictFromDMatrix(self.h, dm.handle, cfg, byref(shape), byref(dim),
                                                              byref(res)))
        n = int(np.prod([shape[i] for i in range(dim.value)]))
        return np.ctypeslib.as_array(res, shape=(n,)).copy()

    def __del__(self):
        try:
            self.lib.XGBoosterFree(self.h)
        except Exception:
            pass


def trees_from_model(model: dict):
    """Per-tree node arrays from the JSON model (split_indices, split_conditions, default_left,
    left/right_children, base_weights, loss_changes, sum_hessian)."""
    out = []
    for t in model["learner"]["gradient_booster"]["model"]["trees"]:
        out.append(dict(
            left=[int(v) for v in t["left_children"]], right=[int(v) for v in t["right_children"]],
            parent=[int(v) for v in t["parents"]],
            findex=[int(v) for v in t["split_indices"]],
            cond=[float(v) for v in t["split_conditions"]],
            default_left=[int(v) for v in t["default_left"]],
            base_weight=[float(v) for v in t["base_weights"]],
            loss_chg=[float(v) for v in t["loss_changes"]],
            sum_hess=[float(v) for v in t["sum_hessian"]]))
    return out


# ----------------------------------------------------------------------------------------------------------------------
# The reference's own distributed plumbing (tracker + communicator, include/xgboost/c_api.h:1669-1853), for the
# 2-process drop-in test of the shim's distributed mode.
# ----------------------------------------------------------------------------------------------------------------------
class Tracker:
    """RabitTracker through XGTrackerCreate / XGTrackerRun / XGTrackerWorkerArgs / XGTrackerWaitFor / XGTrackerFree."""

    def __init__(self, n_workers: int, kind: str = "cuda", host_ip: str = "127.0.0.1", timeout: int = 300):
        self.lib = load(kind)
        self.h = c_void_p()
        cfg = json.dumps({"dmlc_communicator": "rabit", "n_workers": int(n_workers), "host": host_ip, "port": 0,
                          "sortby": 0, "timeout": int(timeout)}).encode()
        _check(self.lib, self.lib.XGTrackerCreate(cfg, byref(self.h)))
        _check(self.lib, self.lib.XGTrackerRun(self.h, b"{}"))

    def worker_args(self) -> dict:
        out = c_char_p()
        _check(self.lib, self.lib.XGTrackerWorkerArgs(self.h, byref(out)))
        return json.loads(out.value.decode())

    def wait(self, timeout: int = 300):
        _check(self.lib, self.lib.XGTrackerWaitFor(self.h, json.dumps({"timeout": int(timeout)}).encode()))

    def free(self):
        if self.h:
            self.lib.XGTrackerFree(self.h)
            self.h = c_void_p()


def communicator_init(kind: str, args: dict) -> None:
    lib = load(kind)
    _check(lib, lib.XGCommunicatorInit(json.dumps(args).encode()))


def communicator_finalize(kind: str) -> None:
    lib = load(kind)
    _check(lib, lib.XGCommunicatorFinalize())


def communicator_rank(kind: str) -> int:
    return int(load(kind).XGCommunicatorGetRank())
