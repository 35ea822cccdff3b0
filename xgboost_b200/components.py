"""Thin Python wrappers over the component entry points of the C ABI (one per SURVEY.md 8(a) row).
Used by the parity tests; each call launches the same kernels the updater launches."""
from __future__ import annotations

from ctypes import POINTER, byref, c_float, c_int32, c_uint32, c_uint64, c_void_p
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import lib as _l
from .updater import EllpackPage, _require_cuda, _stream_ptr


def quantise(gpair: torch.Tensor) -> Tuple[_l.Quantiser, torch.Tensor, Tuple[int, int]]:
    """G1-G3: returns (quantiser, gq int64[n,2] on device, root sum)."""
    _require_cuda(gpair, "gpair")
    n = gpair.shape[0]
    gq = torch.empty((n, 2), dtype=torch.int64, device=gpair.device)
    q = _l.Quantiser()
    root = _l.GPair64()
    _l.check(_l.load().b200hist_quantise(gpair.device.index or 0, c_void_p(gpair.data_ptr()), n, byref(q),
                                       c_void_p(gq.data_ptr()), byref(root), _stream_ptr(gpair.device)))
    return q, gq, (root.grad, root.hess)


def _u64(a: Sequence[int]):
    arr = np.ascontiguousarray(a, dtype=np.uint64)
    return arr, arr.ctypes.data_as(POINTER(c_uint64))


def build_hist(page: EllpackPage, gq: torch.Tensor, ridx: torch.Tensor, seg_begin: Sequence[int],
               seg_count: Sequence[int]) -> torch.Tensor:
    """H2: histograms int64[n_nodes, total_bins, 2] for row segments of ridx."""
    _require_cuda(gq, "gq")
    _require_cuda(ridx, "ridx")
    assert gq.dtype == torch.int64 and ridx.dtype == torch.int32 or ridx.dtype == torch.uint32
    n_nodes = len(seg_begin)
    hist = torch.empty((n_nodes, page.cuts.total_bins, 2), dtype=torch.int64, device=gq.device)
    b, bp = _u64(seg_begin)
    c, cp = _u64(seg_count)
    _l.check(_l.load().b200hist_build_hist(page.handle, c_void_p(gq.data_ptr()), c_void_p(ridx.data_ptr()), bp, cp,
                                         n_nodes, c_void_p(hist.data_ptr()), _stream_ptr(gq.device)))
    return hist


def subtract(parent: torch.Tensor, built: torch.Tensor) -> torch.Tensor:
    """H4: parent - built over int64[total_bins, 2]."""
    _require_cuda(parent, "parent")
    out = torch.empty_like(parent)
    _l.check(_l.load().b200hist_subtract(parent.device.index or 0, c_void_p(parent.data_ptr()),
                                       c_void_p(built.data_ptr()), c_void_p(out.data_ptr()),
                                       parent.numel() // 2, _stream_ptr(parent.device)))
    return out


def evaluate_splits(page: EllpackPage, hist: torch.Tensor, parent_sums: Sequence[Tuple[int, int]],
                    param: _l.TrainParam, quantiser: _l.Quantiser,
                    feature_set: Optional[Sequence[int]] = None) -> List[_l.Candidate]:
    """E1-E4: best split per node; hist int64[n_nodes, total_bins, 2] on device."""
    _require_cuda(hist, "hist")
    n_nodes = len(parent_sums)
    ps = (_l.GPair64 * n_nodes)(*[_l.GPair64(int(g), int(h)) for g, h in parent_sums])
    out = (_l.Candidate * n_nodes)()
    if feature_set is not None:
        fs = np.ascontiguousarray(feature_set, dtype=np.int32)
        fsp, nfs = fs.ctypes.data_as(POINTER(c_int32)), fs.size
    else:
        fsp, nfs = None, 0
    _l.check(_l.load().b200hist_evaluate_splits(page.handle, c_void_p(hist.data_ptr()), ps, n_nodes, fsp, nfs,
                                              byref(param), byref(quantiser), out, _stream_ptr(hist.device)))
    return list(out)


def evaluate_splits_cat(page: EllpackPage, hist: torch.Tensor, parent_sums: Sequence[Tuple[int, int]],
                        param: _l.TrainParam, quantiser: _l.Quantiser, feature_is_cat: Sequence[int],
                        max_cat_to_onehot: int = 4, max_cat_threshold: int = 64, n_cat_words: int = 8,
                        feature_set: Optional[Sequence[int]] = None):
    """E1-E4 with categorical features (b200hist_evaluate_splits_cat): returns (candidates, is_cat flags,
    uint32[n_nodes, n_cat_words] category sets in the reference's LBitField32 layout).  Defaults are TrainParam's
    (max_cat_to_onehot 4, max_cat_threshold 64)."""
    _require_cuda(hist, "hist")
    n_nodes = len(parent_sums)
    ps = (_l.GPair64 * n_nodes)(*[_l.GPair64(int(g), int(h)) for g, h in parent_sums])
    out = (_l.Candidate * n_nodes)()
    if feature_set is not None:
        fs = np.ascontiguousarray(feature_set, dtype=np.int32)
        fsp, nfs = fs.ctypes.data_as(POINTER(c_int32)), fs.size
    else:
        fsp, nfs = None, 0
    fic = np.ascontiguousarray(feature_is_cat, dtype=np.uint8)
    is_cat = np.zeros(n_nodes, dtype=np.int32)
    bits = np.zeros((n_nodes, n_cat_words), dtype=np.uint32)
    _l.check(_l.load().b200hist_evaluate_splits_cat(
        page.handle, c_void_p(hist.data_ptr()), ps, n_nodes, fsp, nfs, byref(param), byref(quantiser),
        c_void_p(fic.ctypes.data), int(max_cat_to_onehot), int(max_cat_threshold), out,
        is_cat.ctypes.data_as(POINTER(c_int32)), bits.ctypes.data_as(POINTER(c_uint32)), int(n_cat_words),
        _stream_ptr(hist.device)))
    return list(out), is_cat, bits


def col_sample_stream(features: Sequence[int], colsample: float, seeds: Sequence[int], device: int = 0) -> np.ndarray:
    """b200hist_col_sample_stream: int32[len(seeds), n_sel], draw d = the reference's ColSample with engine seed
    seeds[d]."""
    f = np.ascontiguousarray(features, dtype=np.int32)
    sd = np.ascontiguousarray([int(s) & 0xFFFFFFFF for s in seeds], dtype=np.uint32)
    n_sel = c_uint32(0)
    out = np.zeros((max(sd.size, 1), max(f.size, 1)), dtype=np.int32).reshape(-1)
    _l.check(_l.load().b200hist_col_sample_stream(int(device), f.ctypes.data_as(POINTER(c_int32)), f.size,
                                                 float(colsample), sd.ctypes.data_as(POINTER(c_uint32)), sd.size,
                                                 out.ctypes.data_as(POINTER(c_int32)), byref(n_sel), c_void_p()))
    return out[: sd.size * n_sel.value].reshape(sd.size, n_sel.value).copy()


def partition(page: EllpackPage, ridx: torch.Tensor, seg_begin: Sequence[int], seg_count: Sequence[int],
              findex: Sequence[int], split_cond: Sequence[float], default_left: Sequence[bool]
              ) -> Tuple[torch.Tensor, List[int]]:
    """P2: stable partition of each segment; returns (new ridx, left counts)."""
    _require_cuda(ridx, "ridx")
    out = ridx.clone()
    n_nodes = len(seg_begin)
    b, bp = _u64(seg_begin)
    c, cp = _u64(seg_count)
    f = np.ascontiguousarray(findex, dtype=np.int32)
    sc = np.ascontiguousarray(split_cond, dtype=np.float32)
    dl = np.ascontiguousarray([1 if d else 0 for d in default_left], dtype=np.int32)
    left = np.zeros(n_nodes, dtype=np.uint64)
    _l.check(_l.load().b200hist_partition(page.handle, c_void_p(ridx.data_ptr()), c_void_p(out.data_ptr()), bp, cp,
                                        f.ctypes.data_as(POINTER(c_int32)), sc.ctypes.data_as(POINTER(c_float)),
                                        dl.ctypes.data_as(POINTER(c_int32)), n_nodes,
                                        left.ctypes.data_as(POINTER(c_uint64)), _stream_ptr(ridx.device)))
    return out, [int(x) for x in left]


def smem_atomic_peak(h_lo_zero: bool = True, device: Optional[torch.device] = None) -> float:
    """Measured ceiling of the histogram kernel's shared-memory accumulation on this device, in (row, feature)
    updates per clock per SM (b200hist_smem_atomic_peak)."""
    from ctypes import c_double
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = c_double(0.0)
    _l.check(_l.load().b200hist_smem_atomic_peak(device.index or 0, 1 if h_lo_zero else 0, byref(out),
                                               _stream_ptr(device)))
    return float(out.value)
