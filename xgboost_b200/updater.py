"""Host-side mirror of the reference interface for the hist tree-updater path.

Names follow the reference (paths relative to /root/reference):
  HistogramCuts   src/common/hist_util.h:38-190
  EllpackPage     src/data/ellpack_page.cuh:181-391
  GPUHistMaker    src/tree/updater_gpu_hist.cu:758-891  (TreeUpdater: Configure / Update /
                  UpdatePredictionCache / Name / HasNodePosition, include/xgboost/tree_updater.h:37-108)
  RegTree         include/xgboost/tree_model.h (only what ExpandNode records)

PyTorch is used only to own device memory and streams; all compute goes through the C ABI
(include/b200hist.h) into hand-written sm_100a kernels.
"""
from __future__ import annotations

import ctypes
import json
import math
from ctypes import POINTER, byref, c_float, c_int32, c_uint32, c_uint64, c_void_p
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import lib as _l


def _stream_ptr(device: torch.device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise _l.B200HistError("%s must be a CUDA tensor (there is no CPU fallback)" % name)


class HistogramCuts:
    """cut_ptrs (uint32[F+1]) + cut_values (float32[total_bins]); bin b of feature f covers values
    below cut_values[cut_ptrs[f] + b] (its upper bound)."""

    def __init__(self, cut_ptrs: Sequence[int], cut_values: Sequence[float]):
        self.cut_ptrs = np.ascontiguousarray(cut_ptrs, dtype=np.uint32)
        self.cut_values = np.ascontiguousarray(cut_values, dtype=np.float32)
        if self.cut_ptrs.ndim != 1 or self.cut_ptrs.size < 2 or self.cut_ptrs[0] != 0:
            raise ValueError("cut_ptrs must be a 1-D array starting at 0")
        if int(self.cut_ptrs[-1]) != self.cut_values.size:
            raise ValueError("cut_ptrs[-1] must equal len(cut_values)")

    @property
    def n_features(self) -> int:
        return int(self.cut_ptrs.size - 1)

    @property
    def total_bins(self) -> int:
        return int(self.cut_values.size)

    @classmethod
    def from_dense(cls, x: torch.Tensor, max_bin: int = 256) -> "HistogramCuts":
        """Quantile sketch of a dense float32 CUDA matrix (NaN = missing): the cuts the reference's GPU sketch
        produces for an in-core, unweighted DMatrix (common::DeviceSketch)."""
        _require_cuda(x, "x")
        if x.dtype != torch.float32 or x.dim() != 2 or (x.shape[1] > 1 and x.stride(1) != 1):
            raise ValueError("x must be a row-major float32 matrix")
        n, f = x.shape
        ptrs = np.zeros(f + 1, dtype=np.uint32)
        vals = np.zeros(f * max_bin, dtype=np.float32)
        _l.check(_l.load().b200hist_sketch_dense(x.device.index or 0, c_void_p(x.data_ptr()), n, x.stride(0), f, max_bin,
                                               ptrs.ctypes.data_as(POINTER(c_uint32)),
                                               vals.ctypes.data_as(POINTER(c_float)), vals.size,
                                               _stream_ptr(x.device)))
        return cls(ptrs, vals[: int(ptrs[-1])].copy())

    @classmethod
    def from_batches(cls, batches, max_bin: int = 256) -> "HistogramCuts":
        """Cuts of QuantileDMatrix(DataIter over `batches`): every batch (dense float32 CUDA matrix, NaN =
        missing) is one sketch window merged into a resident summary (quantile_dmatrix.cu:29-126)."""
        sketch = None
        for x in batches:
            if sketch is None:
                sketch = QuantileSketch(x.shape[1], max_bin, x.device)
            sketch.push(x)
        if sketch is None:
            raise ValueError("no batches")
        return sketch.make_cuts()

    def search_bin(self, value: float, fidx: int) -> int:
        """HistogramCuts::SearchBin (hist_util.h:112-121), host helper for tests."""
        beg, end = int(self.cut_ptrs[fidx]), int(self.cut_ptrs[fidx + 1])
        idx = beg + int(np.searchsorted(self.cut_values[beg:end], np.float32(value), side="right"))
        return idx - 1 if idx == end else idx


class QuantileSketch:
    """Resident per-feature quantile summaries on the device: common::SketchContainer (quantile.cuh:38-200) as
    QuantileDMatrix drives it -- push one DataIter batch at a time, then make_cuts()."""

    def __init__(self, n_features: int, max_bin: int = 256, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_features, self.max_bin = int(n_features), int(max_bin)
        self._h = c_void_p()
        self.fix_error_changes = 0
        _l.check(_l.load().b200hist_sketch_create(self.device.index or 0, self.n_features, self.max_bin, byref(self._h)))

    def set_feature_types(self, feature_is_cat: Optional[Sequence[int]]) -> None:
        """Categorical columns (non-zero flag) get the cut values 0 .. largest code instead of quantiles."""
        self._has_cat = bool(feature_is_cat is not None and any(feature_is_cat))
        if not self._has_cat:
            _l.check(_l.load().b200hist_sketch_set_feature_types(self._h, None, 0))
            return
        flags = np.ascontiguousarray([1 if v else 0 for v in feature_is_cat], dtype=np.uint8)
        _l.check(_l.load().b200hist_sketch_set_feature_types(self._h, c_void_p(flags.ctypes.data), flags.size))

    def push(self, x: torch.Tensor) -> None:
        _require_cuda(x, "x")
        if x.dtype != torch.float32 or x.dim() != 2 or x.shape[1] != self.n_features or \
                (x.shape[1] > 1 and x.stride(1) != 1):
            raise ValueError("x must be a row-major float32 matrix with n_features columns")
        ld = x.stride(0) if x.shape[0] > 1 else max(x.shape[1], 1)
        _l.check(_l.load().b200hist_sketch_push_dense(self._h, c_void_p(x.data_ptr()), x.shape[0], ld,
                                                    _stream_ptr(x.device)))

    def make_cuts(self) -> "HistogramCuts":
        ptrs = np.zeros(self.n_features + 1, dtype=np.uint32)
        # a categorical column may hold more categories than max_bin
        vals = np.zeros(self.n_features * self.max_bin + (self.n_features * 1024 if getattr(self, "_has_cat", False) else 0),
                        dtype=np.float32)
        fixes = c_uint32(0)
        _l.check(_l.load().b200hist_sketch_make_cuts(self._h, ptrs.ctypes.data_as(POINTER(c_uint32)),
                                                   vals.ctypes.data_as(POINTER(c_float)), vals.size, byref(fixes),
                                                   _stream_ptr(self.device)))
        self.fix_error_changes = int(fixes.value)
        return HistogramCuts(ptrs, vals[: int(ptrs[-1])].copy())

    def close(self) -> None:
        if self._h:
            _l.load().b200hist_sketch_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EllpackPage:
    """Quantised feature matrix resident in HBM (our layout, see include/b200hist.h)."""

    def __init__(self, cuts: HistogramCuts, n_rows: int, device: int = 0, has_missing: bool = False,
                 sparse_row_stride: Optional[int] = None, page_rows: Optional[int] = None):
        """page_rows: external memory -- the matrix lives in pinned host memory and the updater streams it through two
        device staging buffers of that many rows (b200hist_ellpack_create_paged); None = resident in HBM."""
        lib = _l.load()
        self.cuts = cuts
        self.device = int(device)
        self.n_rows = int(n_rows)
        self._h = c_void_p()
        cp = cuts.cut_ptrs.ctypes.data_as(POINTER(c_uint32))
        cv = cuts.cut_values.ctypes.data_as(POINTER(c_float))
        if sparse_row_stride is not None:  # sparse ELLPACK layout (global-bin symbols)
            _l.check(lib.b200hist_ellpack_create_sparse(self.device, self.n_rows, cuts.n_features, cp, cv,
                                                        int(sparse_row_stride), byref(self._h)))
        elif page_rows is not None:
            _l.check(lib.b200hist_ellpack_create_paged(self.device, self.n_rows, cuts.n_features, cp, cv,
                                                       1 if has_missing else 0, int(page_rows), byref(self._h)))
        else:
            _l.check(lib.b200hist_ellpack_create(self.device, self.n_rows, cuts.n_features, cp, cv,
                                                 1 if has_missing else 0, byref(self._h)))

    def paging_info(self) -> Dict[str, int]:
        """n_pages (0 = in-core), page_rows and the bytes the updater has streamed host -> device so far."""
        n_pages, rows, h2d = c_uint32(), c_uint64(), c_uint64()
        _l.check(_l.load().b200hist_ellpack_paging_info(self._h, byref(n_pages), byref(rows), byref(h2d)))
        return {"n_pages": int(n_pages.value), "page_rows": int(rows.value), "h2d_bytes": int(h2d.value)}

    def push_csr(self, indptr: torch.Tensor, indices: torch.Tensor, values: torch.Tensor, row_begin: int) -> None:
        """Bin a CSR batch (device tensors: int64 indptr relative to the batch, int32 indices, float32 values)."""
        for t, name in ((indptr, "indptr"), (indices, "indices"), (values, "values")):
            _require_cuda(t, name)
        if indptr.dtype != torch.int64 or indices.dtype != torch.int32 or values.dtype != torch.float32:
            raise ValueError("CSR batch must be int64 indptr, int32 indices, float32 values")
        _l.check(_l.load().b200hist_ellpack_push_csr(self._h, int(row_begin), c_void_p(indptr.data_ptr()),
                                                   c_void_p(indices.data_ptr()), c_void_p(values.data_ptr()),
                                                   indptr.numel() - 1, _stream_ptr(values.device)))

    @classmethod
    def from_csr(cls, indptr, indices, values, cuts: HistogramCuts, device: int = 0) -> "EllpackPage":
        """Sparse page from host CSR arrays (numpy); row_stride = the longest row (GetRowStride)."""
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        n = indptr.size - 1
        keep = ~np.isnan(np.asarray(values, dtype=np.float32))
        counts = np.add.reduceat(keep.astype(np.int64), indptr[:-1]) if n and keep.size else np.zeros(n, dtype=np.int64)
        counts[np.diff(indptr) == 0] = 0
        stride = max(1, int(counts.max()) if n else 1)
        page = cls(cuts, n, device, sparse_row_stride=stride)
        dev = "cuda:%d" % device
        page.push_csr(torch.from_numpy(indptr).to(dev), torch.from_numpy(np.ascontiguousarray(indices, dtype=np.int32)).to(dev),
                      torch.from_numpy(np.ascontiguousarray(values, dtype=np.float32)).to(dev), 0)
        return page

    @property
    def handle(self) -> c_void_p:
        return self._h

    def set_feature_types(self, feature_is_cat: Optional[Sequence[int]]) -> None:
        """Categorical columns (non-zero flag) of later push_dense calls are binned by SearchCatBin: a category's
        bin is its position among the feature's category codes.  Dense-layout pages; call before push_dense."""
        if feature_is_cat is None or not any(feature_is_cat):
            _l.check(_l.load().b200hist_ellpack_set_feature_types(self._h, None, 0))
            return
        flags = np.ascontiguousarray([1 if v else 0 for v in feature_is_cat], dtype=np.uint8)
        _l.check(_l.load().b200hist_ellpack_set_feature_types(self._h, c_void_p(flags.ctypes.data), flags.size))

    def push_dense(self, x: torch.Tensor, row_begin: int) -> None:
        """Bin a dense float32 CUDA batch into rows [row_begin, row_begin + len(x))."""
        _require_cuda(x, "x")
        if x.dtype != torch.float32 or x.dim() != 2 or (x.shape[1] > 1 and x.stride(1) != 1):
            raise ValueError("x must be a row-major float32 matrix")
        if x.shape[1] == 1:
            x = x.reshape(-1).contiguous().reshape(-1, 1)  # normalise the stride of the unit dimension
        _l.check(_l.load().b200hist_ellpack_push_dense(self._h, int(row_begin), c_void_p(x.data_ptr()),
                                                     x.shape[0], x.stride(0), _stream_ptr(x.device)))

    @classmethod
    def from_dense(cls, x: torch.Tensor, cuts: HistogramCuts, has_missing: Optional[bool] = None,
                   feature_is_cat: Optional[Sequence[int]] = None, page_rows: Optional[int] = None) -> "EllpackPage":
        _require_cuda(x, "x")
        if has_missing is None:
            has_missing = bool(torch.isnan(x).any().item())
        page = cls(cuts, x.shape[0], x.device.index or 0, has_missing, page_rows=page_rows)
        if feature_is_cat is not None:
            page.set_feature_types(feature_is_cat)
        page.push_dense(x.contiguous(), 0)
        return page

    @classmethod
    def from_batches(cls, batches: Iterable[torch.Tensor], n_rows: int, cuts: HistogramCuts, device: int = 0,
                     has_missing: bool = False) -> "EllpackPage":
        """QuantileDMatrix-style construction from a device batch iterator
        (src/data/iterative_dmatrix.cu:20-97 pass 2)."""
        page = cls(cuts, n_rows, device, has_missing)
        at = 0
        for b in batches:
            page.push_dense(b, at)
            at += b.shape[0]
        if at != n_rows:
            raise ValueError("batches hold %d rows, expected %d" % (at, n_rows))
        return page

    def import_reference(self, gidx_buffer: torch.Tensor, row_stride: int, n_symbols: int, is_dense: bool) -> None:
        _require_cuda(gidx_buffer, "gidx_buffer")
        _l.check(_l.load().b200hist_ellpack_import_reference(self._h, c_void_p(gidx_buffer.data_ptr()), row_stride,
                                                           n_symbols, 1 if is_dense else 0,
                                                           _stream_ptr(gidx_buffer.device)))

    def info(self) -> Dict[str, int]:
        n_rows, nf, tb, rb, bb, ns = c_uint64(), c_uint32(), c_uint32(), c_uint32(), c_uint32(), c_uint32()
        ptr = c_void_p()
        _l.check(_l.load().b200hist_ellpack_info(self._h, byref(n_rows), byref(nf), byref(tb), byref(rb), byref(bb),
                                               byref(ns), byref(ptr)))
        return dict(n_rows=n_rows.value, n_features=nf.value, total_bins=tb.value, row_bytes=rb.value,
                    bin_bytes=bb.value, null_symbol=ns.value, bins_ptr=ptr.value)

    def symbols(self) -> np.ndarray:
        """Feature-local symbols as an int64 host matrix [n_rows, n_features] (null -> -1); tests only."""
        i = self.info()
        raw = np.zeros((i["n_rows"], i["row_bytes"]), dtype=np.uint8)
        _l.check(_l.load().b200hist_ellpack_copy_bins_to_host(self._h, c_void_p(raw.ctypes.data)))
        if i["bin_bytes"] == 4:  # sparse layout: GLOBAL bins per ELL slot, null -> -1
            sym = raw.view(np.uint32).astype(np.int64)
            sym[sym >= i["total_bins"]] = -1
            return sym
        if i["bin_bytes"] == 1:
            sym = raw[:, : i["n_features"]].astype(np.int64)
        else:
            sym = raw.view(np.uint16)[:, : i["n_features"]].astype(np.int64)
        if i["null_symbol"] != 0xFFFFFFFF:
            sym[sym == i["null_symbol"]] = -1
        return sym

    def __del__(self):
        try:
            if self._h:
                _l.load().b200hist_ellpack_destroy(self._h)
                self._h = c_void_p()
        except Exception:
            pass


class Communicator:
    """NCCL communicator of the per-level histogram allreduce (one process per GPU).

    The 128-byte unique id is created on rank 0 and handed to the other ranks by the caller's control plane
    (`from_torch_distributed` uses a torch.distributed broadcast -- plumbing only; the allreduce itself is
    issued by the C++ tree driver on the kernels' stream)."""

    def __init__(self, device: int, world: int, rank: int, unique_id: bytes):
        if len(unique_id) != 128:
            raise ValueError("unique_id must be 128 bytes")
        self.world, self.rank, self.device = int(world), int(rank), int(device)
        self._h = c_void_p()
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        _l.check(_l.load().b200hist_comm_create(self.device, self.world, self.rank, buf, byref(self._h)))

    @staticmethod
    def unique_id() -> bytes:
        buf = (ctypes.c_uint8 * 128)()
        _l.check(_l.load().b200hist_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_torch_distributed(cls, device: torch.device) -> "Communicator":
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        t = torch.zeros(128, dtype=torch.uint8, device=device if dist.get_backend() == "nccl" else "cpu")
        if rank == 0:
            t.copy_(torch.tensor(list(cls.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        return cls(device.index or 0, world, rank, bytes(t.cpu().tolist()))

    @property
    def handle(self) -> c_void_p:
        return self._h

    def __del__(self):
        try:
            if self._h:
                _l.load().b200hist_comm_destroy(self._h)
                self._h = c_void_p()
        except Exception:
            pass


def shard_groups(n_groups: int, world: int, rank: int):
    """Feature groups [begin, end) owned by `rank` (host logic of the feature-block sharding)."""
    b, e = c_uint32(), c_uint32()
    _l.check(_l.load().b200hist_shard_groups(n_groups, world, rank, byref(b), byref(e)))
    return b.value, e.value


class RegTree:
    """The tree as the updater leaves it (node arrays in ExpandNode order)."""

    _DTYPE = np.dtype([("parent", "<i4"), ("left", "<i4"), ("right", "<i4"), ("findex", "<i4"), ("fvalue", "<f4"),
                       ("default_left", "<i4"), ("is_leaf", "<i4"), ("leaf_value", "<f4"), ("loss_chg", "<f4"),
                       ("sum_hess", "<f4"), ("base_weight", "<f4"), ("depth", "<i4")])

    def __init__(self, nodes, count: Optional[int] = None):
        """nodes: a ctypes array of b200hist_tree_node (the first `count` are taken with one memcpy; the per-node
        dicts are built only when somebody asks) or a sequence of TreeNode."""
        assert ctypes.sizeof(_l.TreeNode) == self._DTYPE.itemsize
        if not isinstance(nodes, ctypes.Array):
            nodes = (_l.TreeNode * len(nodes))(*nodes)
        n = len(nodes) if count is None else int(count)
        raw = ctypes.string_at(ctypes.addressof(nodes), n * self._DTYPE.itemsize)
        self.array = np.frombuffer(raw, dtype=self._DTYPE, count=n).copy()
        self._nodes = None

    @property
    def nodes(self) -> List[dict]:
        if self._nodes is None:
            a = self.array
            self._nodes = [
                dict(nodeid=i, parent=int(a["parent"][i]), left=int(a["left"][i]), right=int(a["right"][i]),
                     findex=int(a["findex"][i]), fvalue=float(a["fvalue"][i]), default_left=bool(a["default_left"][i]),
                     is_leaf=bool(a["is_leaf"][i]), leaf_value=float(a["leaf_value"][i]),
                     loss_chg=float(a["loss_chg"][i]), sum_hess=float(a["sum_hess"][i]),
                     base_weight=float(a["base_weight"][i]), depth=int(a["depth"][i]))
                for i in range(a.shape[0])]
        return self._nodes

    @property
    def num_nodes(self) -> int:
        return int(self.array.shape[0])

    def max_depth(self) -> int:
        return max(n["depth"] for n in self.nodes)

    def dump_json(self) -> dict:
        """Same shape as Booster.get_dump(dump_format='json', with_stats=True) of the reference."""
        def rec(i: int) -> dict:
            n = self.nodes[i]
            if n["is_leaf"]:
                return {"nodeid": i, "leaf": n["leaf_value"], "cover": n["sum_hess"]}
            return {"nodeid": i, "depth": n["depth"], "split": "f%d" % n["findex"], "split_condition": n["fvalue"],
                    "yes": n["left"], "no": n["right"], "missing": n["left"] if n["default_left"] else n["right"],
                    "gain": n["loss_chg"], "cover": n["sum_hess"], "children": [rec(n["left"]), rec(n["right"])]}
        return rec(0)

    def structure(self) -> List[tuple]:
        """(nodeid, findex, fvalue bits, default_left, left, right) per node: the bit-parity signature."""
        out = []
        for n in self.nodes:
            fv = np.float32(n["fvalue"]).view(np.uint32).item() if not n["is_leaf"] else 0
            out.append((n["nodeid"], n["findex"], fv, n["default_left"], n["left"], n["right"]))
        return out


class Forest:
    """Trees resolved against a page's cuts, for prediction on quantised pages (include/b200hist.h, "Prediction on the
    quantised matrix"): the GPUPredictor step for DMatrices bound to the training cuts."""

    def __init__(self, page: EllpackPage):
        self._page = page            # keeps the cuts alive
        self._h = c_void_p()
        _l.check(_l.load().b200hist_forest_create(page.handle, byref(self._h)))

    def add_tree(self, tree: RegTree, is_cat: Optional[Sequence[int]] = None, cat_sets: Optional[Sequence[set]] = None) -> None:
        """tree: what GPUHistMaker.update returned; is_cat / cat_sets: GPUHistMaker.last_tree_categories(...)."""
        a = np.ascontiguousarray(tree.array)
        flags_p, bits_p, words = None, None, 0
        if is_cat is not None and any(is_cat):
            flags = np.ascontiguousarray(is_cat, dtype=np.int32)
            words = max((max(cs) if cs else 0) for cs in cat_sets) // 32 + 1
            bits = np.zeros((a.shape[0], words), dtype=np.uint32)
            for i, cs in enumerate(cat_sets):
                for c in cs:
                    bits[i, c // 32] |= np.uint32(1 << (31 - c % 32))
            flags_p, bits_p = flags.ctypes.data_as(POINTER(c_int32)), bits.ctypes.data_as(POINTER(c_uint32))
        _l.check(_l.load().b200hist_forest_add_tree(self._h, c_void_p(a.ctypes.data), a.shape[0], flags_p, bits_p, words))

    @property
    def num_trees(self) -> int:
        n = c_uint32()
        _l.check(_l.load().b200hist_forest_num_trees(self._h, byref(n)))
        return int(n.value)

    def predict_margin(self, page: EllpackPage, out: torch.Tensor, tree_begin: int = 0, tree_end: Optional[int] = None,
                       row_begin: int = 0) -> torch.Tensor:
        """out[k] += sum of the leaves row row_begin + k reaches in trees [tree_begin, tree_end), in tree order."""
        _require_cuda(out, "out")
        if out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be contiguous float32")
        te = self.num_trees if tree_end is None else int(tree_end)
        _l.check(_l.load().b200hist_forest_predict_margin(self._h, page.handle, int(tree_begin), te, int(row_begin),
                                                        out.numel(), c_void_p(out.data_ptr()),
                                                        _stream_ptr(out.device)))
        return out

    def predict_leaf(self, page: EllpackPage, tree_begin: int = 0, tree_end: Optional[int] = None,
                     as_float: bool = False) -> torch.Tensor:
        """Leaf ids [n_rows, n_trees]; as_float: stored as float32, the layout Predictor::PredictLeaf returns."""
        te = self.num_trees if tree_end is None else int(tree_end)
        out = torch.empty((page.n_rows, te - tree_begin), dtype=torch.float32 if as_float else torch.int32,
                          device=torch.device("cuda", page.device))
        fn = _l.load().b200hist_forest_predict_leaf_f32 if as_float else _l.load().b200hist_forest_predict_leaf
        _l.check(fn(self._h, page.handle, int(tree_begin), te, 0, page.n_rows, c_void_p(out.data_ptr()),
                    _stream_ptr(out.device)))
        return out

    def predict_reference_page(self, gidx_buffer: torch.Tensor, n_rows: int, row_stride: int, n_symbols: int,
                               dense_layout: bool, is_dense: bool, out: torch.Tensor, mode: int = 0, n_groups: int = 1,
                               tree_begin: int = 0, tree_end: Optional[int] = None) -> torch.Tensor:
        """Predict straight from a reference EllpackPage's compressed buffer (uint8 CUDA tensor incl. its 8 bytes of
        padding).  mode 0: out[n_rows] += margin; 1: out[n_rows, n_groups] += per class; 2: out[n_rows, n_trees] = leaf
        ids as floats."""
        _require_cuda(out, "out")
        _require_cuda(gidx_buffer, "gidx_buffer")
        if out.dtype != torch.float32 or not out.is_contiguous() or gidx_buffer.dtype != torch.uint8:
            raise ValueError("out must be contiguous float32, gidx_buffer uint8")
        te = self.num_trees if tree_end is None else int(tree_end)
        _l.check(_l.load().b200hist_forest_predict_reference_page(
            self._h, c_void_p(gidx_buffer.data_ptr()), int(n_rows), int(row_stride), int(n_symbols),
            1 if dense_layout else 0, 1 if is_dense else 0, int(mode), int(tree_begin), te, int(n_groups),
            c_void_p(out.data_ptr()), _stream_ptr(out.device)))
        return out

    def set_tree_groups(self, groups: Sequence[int]) -> None:
        """Output group of every tree added so far (gbm::GBTreeModel::tree_info; multi-class)."""
        g = np.ascontiguousarray(groups, dtype=np.uint32)
        _l.check(_l.load().b200hist_forest_set_tree_groups(self._h, g.ctypes.data_as(POINTER(c_uint32)), g.size))

    def predict_margin_groups(self, page: EllpackPage, out: torch.Tensor, n_groups: int, tree_begin: int = 0,
                              tree_end: Optional[int] = None, row_begin: int = 0) -> torch.Tensor:
        """out[k, g] += the leaves row row_begin + k reaches in the trees of group g, one after the other in tree
        order (the reference's multi-group rule: straight onto the output)."""
        _require_cuda(out, "out")
        if out.dtype != torch.float32 or not out.is_contiguous() or out.numel() % int(n_groups) != 0:
            raise ValueError("out must be contiguous float32 [n_rows, n_groups]")
        te = self.num_trees if tree_end is None else int(tree_end)
        _l.check(_l.load().b200hist_forest_predict_margin_groups(
            self._h, page.handle, int(tree_begin), te, int(n_groups), int(row_begin), out.numel() // int(n_groups),
            c_void_p(out.data_ptr()), _stream_ptr(out.device)))
        return out

    def close(self) -> None:
        if self._h:
            _l.load().b200hist_forest_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GPUHistMaker:
    """TreeUpdater "grow_gpu_hist" served by libb200hist.so."""

    def __init__(self, page: EllpackPage):
        self._page = page
        self._h = c_void_p()
        self._param = _l.TrainParam.make()
        self._max_nodes = 0
        self._nodes_buf = None
        self._last_position = None
        self.device = torch.device("cuda:%d" % page.device)

    # TreeUpdater::Configure(Args) -- unknown keys pass through like UpdateAllowUnknown
    def configure(self, args: Dict[str, object]) -> None:
        known = {f[0] for f in _l.TrainParam._fields_} | {"eta", "lambda", "alpha", "gamma"}
        kw = {k: v for k, v in args.items() if k in known}
        self._param = _l.TrainParam.make(**kw)
        self._destroy()
        lib = _l.load()
        _l.check(lib.b200hist_updater_create(self._page.handle, byref(self._param), byref(self._h)))
        mn = c_int32()
        _l.check(lib.b200hist_updater_max_nodes(self._h, byref(mn)))
        self._max_nodes = mn.value
        self._nodes_buf = (_l.TreeNode * self._max_nodes)()

    def name(self) -> str:
        return "grow_gpu_hist"

    def has_node_position(self) -> bool:
        return True

    def set_comm(self, comm: Optional["Communicator"]) -> None:
        """Attach (or detach with None) the allreduce communicator: this rank then builds only its
        feature-group shard of every histogram (collective::Allreduce in the reference)."""
        self._comm = comm  # keep it alive
        _l.check(_l.load().b200hist_updater_set_comm(self._h, comm.handle if comm is not None else c_void_p()))

    def enable_peer_exchange(self, row_split: bool = False) -> None:
        """Map every rank's histogram cache over NVLink P2P (IPC handles all-gathered through
        torch.distributed -- control plane only) and switch the per-level collective from NCCL to the
        peer-memory push + flag barrier kernel.  row_split=True: this rank's page holds ITS slice of the rows
        (same cuts everywhere) and the per-level kernel sums full-width partial histograms instead."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        buf = (ctypes.c_uint8 * 64)()
        _l.check(_l.load().b200hist_updater_ipc_handle(self._h, buf))
        mine = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=self.device)
        allh = [torch.zeros(64, dtype=torch.uint8, device=self.device) for _ in range(world)]
        dist.all_gather(allh, mine)
        flat = (ctypes.c_uint8 * (64 * world)).from_buffer_copy(b"".join(bytes(t.cpu().tolist()) for t in allh))
        _l.check(_l.load().b200hist_updater_set_peers(self._h, rank, world, flat))
        if row_split:
            _l.check(_l.load().b200hist_updater_set_row_split(self._h, 1))
        dist.barrier()

    def set_tree_sampling(self, feature_set: Optional[Sequence[int]] = None, subsample: float = 1.0,
                          seed: int = 0) -> None:
        """Sampling of the next trees: `feature_set` = sorted feature ids of the colsample_bytree draw (None = all),
        `subsample` + `seed` = uniform row sampling with the tree's ctx->Rng()() draw (sampler.cu:160-173)."""
        if feature_set is not None:
            fs = np.ascontiguousarray(feature_set, dtype=np.int32)
            _l.check(_l.load().b200hist_updater_set_tree_sampling(self._h, fs.ctypes.data_as(POINTER(c_int32)), fs.size,
                                                                float(subsample), int(seed) & 0xFFFFFFFF))
        else:
            _l.check(_l.load().b200hist_updater_set_tree_sampling(self._h, None, 0, float(subsample),
                                                                int(seed) & 0xFFFFFFFF))

    def set_level_feature_sets(self, sets) -> None:
        """colsample_bylevel: `sets[d]` = sorted feature ids for nodes at depth d (the last row serves deeper nodes);
        None removes them."""
        if sets is None:
            _l.check(_l.load().b200hist_updater_set_level_feature_sets(self._h, None, 0, 0))
            return
        a = np.ascontiguousarray(sets, dtype=np.int32)
        if a.ndim != 2:
            raise ValueError("sets must be [n_levels][n_per_level]")
        _l.check(_l.load().b200hist_updater_set_level_feature_sets(self._h, a.ctypes.data_as(POINTER(c_int32)),
                                                                 a.shape[0], a.shape[1]))

    def last_max_eval_depth(self) -> int:
        d = c_int32(0)
        _l.check(_l.load().b200hist_updater_last_max_eval_depth(self._h, byref(d)))
        return int(d.value)

    def set_monotone_constraints(self, constraints: Optional[Sequence[int]]) -> None:
        """TrainParam::monotone_constraints: sign per feature (+1 increasing, -1 decreasing, 0 free); None/[] = none."""
        if constraints is None or len(constraints) == 0:
            _l.check(_l.load().b200hist_updater_set_monotone_constraints(self._h, None, 0))
        else:
            c = np.ascontiguousarray(constraints, dtype=np.int32)
            _l.check(_l.load().b200hist_updater_set_monotone_constraints(self._h, c.ctypes.data_as(POINTER(c_int32)), c.size))

    def set_column_sampler(self, tree_set, colsample_bylevel: float, colsample_bynode: float, seeds) -> None:
        """The tree's column sampler on the device: `tree_set` (colsample_bytree draw), the two fractions, and the engine
        outputs the sampler would consume, in order (>= max_nodes + max depth + 1).  None for `seeds` removes it."""
        if seeds is None:
            _l.check(_l.load().b200hist_updater_set_column_sampler(self._h, None, 0, 1.0, 1.0, None, 0))
            return
        ts = np.ascontiguousarray(tree_set, dtype=np.int32)
        sd = np.ascontiguousarray([int(v) & 0xFFFFFFFF for v in seeds], dtype=np.uint32)
        _l.check(_l.load().b200hist_updater_set_column_sampler(self._h, ts.ctypes.data_as(POINTER(c_int32)), ts.size,
                                                             float(colsample_bylevel), float(colsample_bynode),
                                                             sd.ctypes.data_as(POINTER(c_uint32)), sd.size))

    def last_draw_count(self) -> int:
        n = c_int32()
        _l.check(_l.load().b200hist_updater_last_draw_count(self._h, byref(n)))
        return int(n.value)

    def set_sampling_method(self, method: str) -> None:
        """TrainParam::sampling_method: "uniform" or "gradient_based" (applies when subsample < 1 in
        set_tree_sampling).  gradient_based has not run on hardware -- see include/b200hist.h."""
        if method not in ("uniform", "gradient_based"):
            raise ValueError("sampling_method must be 'uniform' or 'gradient_based'")
        _l.check(_l.load().b200hist_updater_set_sampling_method(self._h, 1 if method == "gradient_based" else 0))

    @property
    def max_nodes(self) -> int:
        """Node capacity of one tree (TrainParam::MaxNodes)."""
        return int(self._max_nodes)

    def set_node_feature_sets(self, sets) -> None:
        """colsample_bynode: int32[n_sets >= max_nodes, n_per_set], the tree's stream of per-node draws (the k-th node
        evaluated takes row k); None removes it.  Not yet run on hardware -- see include/b200hist.h."""
        if sets is None:
            _l.check(_l.load().b200hist_updater_set_node_feature_sets(self._h, None, 0, 0))
            return
        a = np.ascontiguousarray(sets, dtype=np.int32)
        if a.ndim != 2:
            raise ValueError("sets must be [n_sets, n_per_set]")
        _l.check(_l.load().b200hist_updater_set_node_feature_sets(self._h, a.ctypes.data_as(POINTER(c_int32)),
                                                                a.shape[0], a.shape[1]))

    def last_eval_count(self) -> int:
        n = c_int32()
        _l.check(_l.load().b200hist_updater_last_eval_count(self._h, byref(n)))
        return int(n.value)

    def set_categorical(self, feature_is_cat: Optional[Sequence[int]], max_cat_to_onehot: int = 4,
                        max_cat_threshold: int = 64) -> None:
        """MetaInfo::feature_types / TrainParam::max_cat_to_onehot, max_cat_threshold: flags per feature (non-zero =
        categorical; its cut values must be its category codes).  None / all zero = numerical only.
        Not yet run on hardware -- see the STATUS note in include/b200hist.h."""
        if feature_is_cat is None or not any(feature_is_cat):
            _l.check(_l.load().b200hist_updater_set_categorical(self._h, None, 0, 4, 64))
            self._has_cat = False
            return
        flags = np.ascontiguousarray([1 if v else 0 for v in feature_is_cat], dtype=np.uint8)
        _l.check(_l.load().b200hist_updater_set_categorical(self._h, c_void_p(flags.ctypes.data), flags.size,
                                                          int(max_cat_to_onehot), int(max_cat_threshold)))
        self._has_cat = True

    def last_tree_categories(self, n_nodes: int):
        """(is_cat[n_nodes], list of category sets) of the last tree: a categorical split sends the rows whose
        category is IN the set to the right child (RegTree::ExpandCategorical's split_cat)."""
        w = c_uint32()
        _l.check(_l.load().b200hist_updater_cat_words(self._h, byref(w)))
        is_cat = np.zeros(max(n_nodes, 1), dtype=np.int32)
        bits = np.zeros((max(n_nodes, 1), max(w.value, 1)), dtype=np.uint32)
        _l.check(_l.load().b200hist_updater_last_tree_categories(self._h, is_cat.ctypes.data_as(POINTER(c_int32)),
                                                               bits.ctypes.data_as(POINTER(c_uint32))))
        sets = [{32 * wi + (31 - b) for wi in range(bits.shape[1]) for b in range(32) if (int(bits[i, wi]) >> b) & 1}
                if is_cat[i] else set() for i in range(n_nodes)]
        return is_cat[:n_nodes], sets

    def set_interaction_constraints(self, sets: Optional[Sequence[Sequence[int]]]) -> None:
        """TrainParam::interaction_constraints, parsed: e.g. [[0, 1], [2, 3, 4]]; None/[] = none.
        (FeatureInteractionConstraintDevice, src/tree/constraints.cu; see include/b200hist.h for the reading of Query
        that is reproduced.)  Not yet run on hardware -- see the STATUS note in include/b200hist.h."""
        if sets is None or len(sets) == 0:
            _l.check(_l.load().b200hist_updater_set_interaction_constraints(self._h, None, None, 0))
            return
        ptr = np.zeros(len(sets) + 1, dtype=np.int32)
        ptr[1:] = np.cumsum([len(s) for s in sets])
        feat = np.ascontiguousarray([f for s in sets for f in s] or [0], dtype=np.int32)
        _l.check(_l.load().b200hist_updater_set_interaction_constraints(
            self._h, ptr.ctypes.data_as(POINTER(c_int32)), feat.ctypes.data_as(POINTER(c_int32)), len(sets)))

    def disable_peer_exchange(self) -> None:
        """Unmap the peers' histogram caches and meet the other ranks at a barrier, so that nobody frees memory that
        a peer still maps; call on every rank before dropping an updater that used enable_peer_exchange()."""
        import torch.distributed as dist
        _l.check(_l.load().b200hist_updater_clear_peers(self._h))
        if dist.is_available() and dist.is_initialized():
            dist.barrier()

    def num_feature_groups(self) -> int:
        n = c_uint32()
        _l.check(_l.load().b200hist_updater_num_feature_groups(self._h, byref(n)))
        return n.value

    def set_profiling(self, on: bool) -> None:
        _l.check(_l.load().b200hist_updater_set_profiling(self._h, 1 if on else 0))

    def last_profile(self) -> Dict[str, float]:
        ms = (c_float * 7)()
        rv = c_uint64()
        _l.check(_l.load().b200hist_updater_last_profile(self._h, ms, byref(rv)))
        names = ["quantise", "hist", "subtract", "evaluate", "apply", "partition", "finalise"]
        d = {k: float(v) for k, v in zip(names, ms)}
        d["row_visits"] = int(rv.value)
        comm = c_float(0.0)
        _l.check(_l.load().b200hist_updater_last_comm_ms(self._h, byref(comm)))
        d["hist_exchange"] = float(comm.value)   # part of "hist"
        n = c_int32()
        buf = (c_float * 64)()
        _l.check(_l.load().b200hist_updater_last_hist_steps(self._h, buf, 64, byref(n)))
        if 0 < n.value <= 64:
            d["hist_levels"] = [float(buf[i]) for i in range(n.value)]
        return d

    # TreeUpdater::Update for one tree
    def update(self, gpair: torch.Tensor, out_position: Optional[torch.Tensor] = None) -> RegTree:
        if not self._h:
            raise _l.B200HistError("configure() must be called before update()")
        n_nodes = c_int32()
        pos_ptr = c_void_p(out_position.data_ptr()) if out_position is not None else c_void_p()
        self._last_position = out_position  # kept alive for update_prediction_cache (None = the updater's own buffer)
        if out_position is not None:
            _require_cuda(out_position, "out_position")
            if out_position.dtype != torch.int32 or out_position.numel() != self._page.n_rows:
                raise ValueError("out_position must be int32[n_rows]")
        if gpair.dtype != torch.float32 or gpair.numel() != 2 * self._page.n_rows or not gpair.is_contiguous():
            raise ValueError("gpair must be contiguous float32[n_rows, 2]")
        if gpair.is_cuda:
            _l.check(_l.load().b200hist_updater_update_tree(self._h, c_void_p(gpair.data_ptr()),
                                                          _stream_ptr(self.device), self._nodes_buf, byref(n_nodes),
                                                          pos_ptr))
        else:  # host gradient: the end-to-end entry (H2D of gpair + D2H of the tree inside)
            _l.check(_l.load().b200hist_updater_update_tree_host(self._h, c_void_p(gpair.data_ptr()),
                                                               _stream_ptr(self.device), self._nodes_buf,
                                                               byref(n_nodes), pos_ptr))
        return RegTree(self._nodes_buf, n_nodes.value)

    def update_async(self, gpair: torch.Tensor, out_position: Optional[torch.Tensor] = None) -> None:
        """update() without waiting for the tree: kernels and the read-back are enqueued, fetch_tree() returns the trees
        in order (at most two in flight).  update_prediction_cache() may follow at once: it reads the device-side tree."""
        if not self._h:
            raise _l.B200HistError("configure() must be called before update_async()")
        _require_cuda(gpair, "gpair")
        if gpair.dtype != torch.float32 or gpair.numel() != 2 * self._page.n_rows or not gpair.is_contiguous():
            raise ValueError("gpair must be contiguous float32[n_rows, 2]")
        if out_position is not None:
            _require_cuda(out_position, "out_position")
            if out_position.dtype != torch.int32 or out_position.numel() != self._page.n_rows:
                raise ValueError("out_position must be int32[n_rows]")
        self._last_position = out_position
        _l.check(_l.load().b200hist_updater_update_tree_async(
            self._h, c_void_p(gpair.data_ptr()), _stream_ptr(self.device),
            c_void_p(out_position.data_ptr()) if out_position is not None else c_void_p()))

    def fetch_tree(self) -> RegTree:
        n_nodes = c_int32()
        _l.check(_l.load().b200hist_updater_fetch_tree(self._h, self._nodes_buf, byref(n_nodes)))
        return RegTree(self._nodes_buf, n_nodes.value)

    # TreeUpdater::UpdatePredictionCache
    def update_prediction_cache(self, preds: torch.Tensor, leaf_values=None) -> bool:
        """preds: float32[n_rows] or one column of a float32[n_rows, n_groups] matrix (multi-class: the column of the
        class this tree belongs to, GBTree::DoBoost's out.Slice(All, gid)).  leaf_values: leaf value per node id when
        the objective rewrote the tree's leaves after update() (adaptive objectives); default = the updater's own."""
        _require_cuda(preds, "preds")
        if preds.dtype != torch.float32 or preds.numel() != self._page.n_rows or preds.dim() != 1:
            raise ValueError("preds must be float32[n_rows] (a strided column view is fine)")
        stride = preds.stride(0) if preds.numel() > 1 else 1
        if stride < 1:
            raise ValueError("preds must have a positive stride")
        # the positions of the last update(): the tensor handed to update() is kept alive here, so the library never
        # reads a buffer the caller has dropped
        pos = self._last_position
        lv_ptr, n_lv = c_void_p(), 0
        if leaf_values is not None:
            lv = np.ascontiguousarray(leaf_values, dtype=np.float32)
            lv_ptr, n_lv = c_void_p(lv.ctypes.data), int(lv.size)
        _l.check(_l.load().b200hist_updater_update_prediction_cache_ex(
            self._h, c_void_p(preds.data_ptr()), stride, c_void_p(pos.data_ptr()) if pos is not None else c_void_p(),
            lv_ptr, n_lv, _stream_ptr(self.device)))
        return True

    # introspection for parity tests
    def last_quantiser(self) -> _l.Quantiser:
        q = _l.Quantiser()
        _l.check(_l.load().b200hist_updater_last_quantiser(self._h, byref(q)))
        return q

    def node_histogram(self, nidx: int) -> np.ndarray:
        out = np.zeros((self._page.cuts.total_bins, 2), dtype=np.int64)
        _l.check(_l.load().b200hist_updater_node_histogram(self._h, nidx, c_void_p(out.ctypes.data)))
        return out

    def node_candidate(self, nidx: int) -> _l.Candidate:
        c = _l.Candidate()
        _l.check(_l.load().b200hist_updater_node_candidate(self._h, nidx, byref(c)))
        return c

    def _destroy(self):
        if self._h:
            _l.load().b200hist_updater_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass
