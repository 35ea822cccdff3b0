"""CPU restatement of the reference GPU quantile sketch -> HistogramCuts, unweighted, numerical features:
the in-core single-window case (SimpleDMatrix on the device) and the batch-by-batch QuantileDMatrix case
(one window per DataIter batch, merged into a resident summary).

TEST INFRASTRUCTURE ONLY (see oracle/hist_oracle.c header).  Follows, relative to /root/reference:
  DeviceSketchWithHessian / ProcessWeightedBatch     src/common/hist_util.cu:234-397
  RequiredSampleCutsPerColumn                        src/common/hist_util.cu:31-35
  SketchEpsilon, WQuantileSketch::LimitSizeLevel     src/common/quantile.h:579-601,734-738  (kFactor = 8)
  SketchContainer::Push -> PruneImpl                 src/common/quantile.cu:108-153,340-378
  BinarySearchQueryIndex                             src/common/quantile.cu:36-61
  SketchContainer::ScanInput (dedupe)                src/common/quantile.cu:380-418
  SketchContainer::MakeCuts -> WQSummary::QueryCutValues   src/common/quantile.cu:658-709, quantile.h:280-340
  SketchContainer::{Merge,FixError,Prune}, MergeImpl,
  MergePartition, SelectPruneIndices                 src/common/quantile.cu:63-96,180-198,207-335,420-566
  SketchContainer::IntermediateNumCuts               src/common/quantile.cuh:68-72
  AdapterDeviceSketch (one window per batch)         src/common/hist_util.cuh:399-443
  cuda_impl::MakeSketches (QuantileDMatrix)          src/data/quantile_dmatrix.cu:29-126
Pinned by the cut values the reference itself produced on a B200 (tests/golden/ref_gpu_trees.json,
tests/golden/ref_quantile_dmatrix_cuts.json).
Ranks are float32 on the device (exact below 2^24 rows per column); the final rank queries are float64.
"""
from __future__ import annotations

import math

import numpy as np

K_FACTOR = 8


def sketch_epsilon(max_bin: int, n: int) -> float:
    n = max(1, n)
    return 1.0 / float(min(max_bin, n))


def limit_size_level(maxn: int, eps: float) -> int:
    if maxn == 0:
        return 1
    internal = eps / K_FACTOR
    nlevel = 1
    while True:
        limit = int(math.ceil(nlevel / internal)) + 1
        limit = min(maxn, limit)
        if (1 << nlevel) * limit >= maxn:
            return limit
        nlevel += 1


def required_cuts_per_column(max_bin: int, n_rows: int) -> int:
    return min(limit_size_level(n_rows, sketch_epsilon(max_bin, n_rows)), n_rows)


def _query_index(n: int, rank: np.float32) -> int:
    """BinarySearchQueryIndex over the unweighted entries e_i = {rmin=i, rmax=i+1, wmin=1} (float32 math)."""
    f = np.float32
    rank2 = f(rank * f(2))
    if rank2 < f(0) + f(1):
        return 0
    if rank2 >= f(n - 1) + f(n):
        return n - 1
    # upper_bound over (rmin+rmax) = 2i+1 for i in [1, n-2], then step back one
    # i = (number of i in [1, n-2] with 2i+1 <= rank2) -> first element greater, minus 1
    lo, hi = 1, n - 1  # search range [1, n-1)
    while lo < hi:
        mid = (lo + hi) // 2
        if not (rank2 < f(f(mid) + f(mid + 1))):
            lo = mid + 1
        else:
            hi = mid
    i = lo - 1
    # RMinNext(i) + RMaxPrev(i+1) = (i + 1) + ((i + 2) - 1)
    if rank2 < f(f(f(i) + f(1)) + f(f(i + 2) - f(1))):
        return i
    return i + 1


def prune_select(n: int, to: int) -> np.ndarray:
    """Sample indices PruneImpl keeps for a column of n sorted values and a budget of `to` entries."""
    if n <= to:
        return np.arange(n, dtype=np.int64)
    f = np.float32
    out = np.empty(to, dtype=np.int64)
    out[0] = 0
    out[to - 1] = n - 1
    front_rmax = f(1)
    w = f(f(n - 1) - front_rmax)  # back.rmin - front.rmax
    for idx in range(1, to - 1):
        q = f(f(f(idx) * w) / f(f(to) - f(1))) + front_rmax
        out[idx] = _query_index(n, f(q))
    return out


def column_summary(sorted_vals: np.ndarray, to: int):
    """Entries (rmin, rmax, wmin, value) of one column after PruneImpl + ScanInput."""
    n = sorted_vals.size
    if n == 0:
        return np.zeros((0, 4), dtype=np.float32)
    sel = prune_select(n, to)
    rmin = sel.astype(np.float32)
    rmax = (sel + 1).astype(np.float32)
    wmin = np.ones(sel.size, dtype=np.float32)
    val = sorted_vals[sel].astype(np.float32)
    # ScanInput: reverse inclusive scan accumulating "first type" duplicates (same value, different rmin)
    m = sel.size
    o_rmin, o_rmax, o_wmin = rmin.copy(), rmax.copy(), wmin.copy()
    for i in range(m - 2, -1, -1):
        if val[i] - val[i + 1] == 0 and o_rmin[i] - o_rmin[i + 1] != 0:
            w = np.float32(o_wmin[i] + o_wmin[i + 1])
            o_wmin[i] = w
            o_rmax[i] = np.float32(o_rmin[i] + w)
    # SegmentedUnique keeps the first entry of every equal-value run
    keep = np.ones(m, dtype=bool)
    keep[1:] = val[1:] != val[:-1]
    return np.stack([o_rmin[keep], o_rmax[keep], o_wmin[keep], val[keep]], axis=1)


def query_cut_values(summary: np.ndarray, max_bin: int) -> np.ndarray:
    """WQSummary::QueryCutValues (quantile.h:280-340)."""
    n = summary.shape[0]
    if n == 0:
        return np.array([1e-5], dtype=np.float32)
    rmin, rmax, wmin, value = (summary[:, k] for k in range(4))
    cuts = []

    def advance(cursor, v):
        while cursor < n and value[cursor] <= v:
            cursor += 1
        return cursor

    last_cut = value[0]
    nxt = advance(1, last_cut)
    if n <= max_bin:
        while nxt < n:
            cpt = value[nxt]
            cuts.append(cpt)
            last_cut = cpt
            nxt = advance(nxt + 1, last_cut)
    else:
        total = float(rmax[n - 1])
        qc = 0
        for i in range(1, max_bin):
            rank = float(i) * total / float(max_bin)
            rank2 = 2.0 * rank
            while qc < n - 2 and rank2 >= float(np.float32(rmin[qc + 1] + rmax[qc + 1])):
                qc += 1
            rmin_next = np.float32(rmin[qc] + wmin[qc])          # RMinNext
            rmax_prev = np.float32(rmax[qc + 1] - wmin[qc + 1])  # RMaxPrev
            pick = qc if rank2 < float(np.float32(rmin_next + rmax_prev)) else qc + 1
            cpt = value[pick]
            if cpt <= last_cut:
                nxt = advance(nxt, last_cut)
                if nxt == n:
                    break
                cpt = value[nxt]
            elif nxt < n and value[nxt] <= cpt:
                nxt = advance(nxt + 1, cpt)
            cuts.append(cpt)
            last_cut = cpt
    cpt = np.float32(value[n - 1])
    cuts.append(np.float32(cpt + np.float32(np.float32(abs(cpt)) + np.float32(1e-5))))
    return np.array(cuts, dtype=np.float32)


def device_sketch_cuts(x: np.ndarray, max_bin: int):
    """HistogramCuts of a dense float32 matrix with NaN = missing."""
    x = np.asarray(x, dtype=np.float32)
    n_rows, n_cols = x.shape
    per_col = required_cuts_per_column(max_bin, n_rows)
    ptrs, vals = [0], []
    for j in range(n_cols):
        col = x[:, j]
        col = np.sort(col[~np.isnan(col)], kind="stable")
        to = min(col.size, per_col)
        cuts = query_cut_values(column_summary(col, to), max_bin)
        vals.append(cuts)
        ptrs.append(ptrs[-1] + cuts.size)
    return np.array(ptrs, dtype=np.uint32), np.concatenate(vals).astype(np.float32)


# ---------------------------------------------------------------------------------------------------
# Multi-window path: resident summary + Merge / FixError / Prune on general entries
# ---------------------------------------------------------------------------------------------------
F32 = np.float32


def _unique_by_value(col: np.ndarray) -> np.ndarray:
    """dh::SegmentedUnique with detail::SketchUnique: first entry of every run of equal values."""
    if col.shape[0] < 2:
        return col
    keep = np.ones(col.shape[0], dtype=bool)
    keep[1:] = (col[1:, 3] - col[:-1, 3]) != 0
    return col[keep]


def merge_columns(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """MergeImpl for one column (quantile.cu:207-335): x = resident summary, y = incoming one."""
    m, n = x.shape[0], y.shape[0]
    if m == 0:
        return y.copy()
    if n == 0:
        return x.copy()
    xv, yv = x[:, 3], y[:, 3]
    # MergePartition: x entries sort ahead of equal y entries
    pos_x = np.arange(m) + np.searchsorted(yv, xv, side="left")
    from_x = np.zeros(m + n, dtype=bool)
    from_x[pos_x] = True
    a_ind = np.concatenate([[0], np.cumsum(from_x)[:-1]]).astype(np.int64)
    b_ind = np.arange(m + n, dtype=np.int64) - a_ind

    def other_rmin(col, ind):          # RMinNext of the last entry of `col` before the insertion point
        prev = np.maximum(ind - 1, 0)
        v = (col[prev, 0] + col[prev, 2]).astype(F32)
        return np.where(ind == 0, F32(0), v).astype(F32)

    def other_rmax(col, ind):          # RMaxPrev of the entry at the insertion point (rmax of the last at the end)
        size = col.shape[0]
        cur = np.minimum(ind, size - 1)
        v = (col[cur, 1] - col[cur, 2]).astype(F32)
        return np.where(ind == size, col[size - 1, 1], v).astype(F32)

    a_c = np.minimum(a_ind, m - 1)
    b_c = np.minimum(b_ind, n - 1)
    xe, ye = x[a_c], y[b_c]
    fx = np.stack([(xe[:, 0] + other_rmin(y, b_ind)).astype(F32), (xe[:, 1] + other_rmax(y, b_ind)).astype(F32),
                   xe[:, 2], xe[:, 3]], axis=1)
    fy = np.stack([(other_rmin(x, a_ind) + ye[:, 0]).astype(F32), (other_rmax(x, a_ind) + ye[:, 1]).astype(F32),
                   ye[:, 2], ye[:, 3]], axis=1)
    eq = np.stack([(xe[:, 0] + ye[:, 0]).astype(F32), (xe[:, 1] + ye[:, 1]).astype(F32),
                   (xe[:, 2] + ye[:, 2]).astype(F32), xe[:, 3]], axis=1)
    x_done, y_done = a_ind == m, b_ind == n
    take_y = x_done | (~y_done & (xe[:, 3] > ye[:, 3]))
    take_eq = ~x_done & ~y_done & (xe[:, 3] == ye[:, 3])
    out = np.where(take_y[:, None], fy, fx)
    out = np.where(take_eq[:, None], eq, out)
    return out.astype(F32)


def fix_error(col: np.ndarray):
    """SketchContainer::FixError (quantile.cu:548-566).  The reference updates entries in place from
    concurrent threads, so a fix that cascades is scheduling dependent; this restatement reads the
    neighbour as it was before the pass and reports how many entries changed (0 on all pinned data)."""
    if col.shape[0] == 0:
        return col, 0
    out = col.copy()
    prev_rmin = np.concatenate([[F32(0)], col[:-1, 0]]).astype(F32)
    prev_rmax = np.concatenate([[F32(0)], col[:-1, 1]]).astype(F32)
    out[:, 0] = np.maximum(out[:, 0], prev_rmin)
    out[:, 1] = np.maximum(out[:, 1], prev_rmax)
    rmin_next = (out[:, 0] + out[:, 2]).astype(F32)
    out[:, 1] = np.maximum(out[:, 1], rmin_next)
    return out, int(np.any(out != col, axis=1).sum())


def _query_index_general(col: np.ndarray, rank: np.ndarray) -> np.ndarray:
    """BinarySearchQueryIndex (quantile.cu:36-61) over general entries, vectorised over ranks (float32)."""
    n = col.shape[0]
    s = (col[:, 0] + col[:, 1]).astype(F32)
    rank2 = (rank * F32(2)).astype(F32)
    i = np.searchsorted(s[1:n - 1], rank2, side="right").astype(np.int64)   # upper_bound over [1, n-1) then -1
    i_c = np.minimum(i, n - 2)
    bound = ((col[i_c, 0] + col[i_c, 2]).astype(F32) + (col[i_c + 1, 1] - col[i_c + 1, 2]).astype(F32)).astype(F32)
    res = np.where(rank2 < bound, i_c, i_c + 1)
    res = np.where(rank2 >= s[n - 1], n - 1, res)
    res = np.where(rank2 < s[0], 0, res)
    return res.astype(np.int64)


def prune_column(col: np.ndarray, to: int) -> np.ndarray:
    """SelectPruneIndices + SegmentedUnique(indices) + gather + SegmentedUnique(values) for one column
    (quantile.cu:63-96,420-489)."""
    n = col.shape[0]
    if n <= to:
        return col
    sel = np.empty(to, dtype=np.int64)
    sel[0] = 0
    sel[to - 1] = n - 1
    if to > 2:
        idx = np.arange(1, to - 1, dtype=np.float32)
        w = F32(col[n - 1, 0] - col[0, 1])                      # back.rmin - front.rmax
        q = ((idx * w).astype(F32) / F32(F32(to) - F32(1))).astype(F32) + F32(col[0, 1])
        sel[1:to - 1] = _query_index_general(col, q.astype(F32))
    keep = np.ones(to, dtype=bool)
    keep[1:] = sel[1:] != sel[:-1]
    return _unique_by_value(col[sel[keep]])


class SketchContainer:
    """Resident per-column summaries (quantile.cuh:38-200); columns hold float32 rows (rmin, rmax, wmin, value)."""

    def __init__(self, n_columns: int, max_bin: int):
        self.max_bin = int(max_bin)
        self.columns = [np.zeros((0, 4), dtype=F32) for _ in range(n_columns)]
        self.rows_seen = 0
        self.fix_error_changes = 0

    def intermediate_num_cuts(self) -> int:
        rows = max(1, self.rows_seen)
        return limit_size_level(rows, sketch_epsilon(self.max_bin, rows)) * len(self.columns)

    def push(self, sorted_columns, num_cuts: int, n_rows_in_batch: int):
        """SketchContainer::Push (quantile.cu:340-378): sorted_columns[j] = ascending non-missing values."""
        self.rows_seen += int(n_rows_in_batch)
        incoming = [column_summary(c, min(c.size, num_cuts)) for c in sorted_columns]
        if sum(s.shape[0] for s in incoming) == 0:
            return
        for j, s in enumerate(incoming):                     # Merge + normalize_merged
            merged = _unique_by_value(merge_columns(self.columns[j], s))
            merged, changed = fix_error(merged)
            self.fix_error_changes += changed
            self.columns[j] = merged
        self.prune(self.intermediate_num_cuts())

    def prune(self, to: int):
        """SketchContainer::Prune (quantile.cu:420-489); columns no longer than `to` pass through."""
        self.columns = [prune_column(c, int(to)) for c in self.columns]

    def make_cuts(self):
        ptrs, vals = [0], []
        for c in self.columns:
            cuts = query_cut_values(c, self.max_bin)
            vals.append(cuts)
            ptrs.append(ptrs[-1] + cuts.size)
        return np.array(ptrs, dtype=np.uint32), np.concatenate(vals).astype(np.float32)


def _sorted_columns(x: np.ndarray):
    x = np.asarray(x, dtype=np.float32)
    return [np.sort(x[:, j][~np.isnan(x[:, j])], kind="stable") for j in range(x.shape[1])]


def quantile_dmatrix_cuts(batches, max_bin: int):
    """HistogramCuts of QuantileDMatrix(DataIter over `batches`) on the device (quantile_dmatrix.cu:29-126):
    every batch is one sketch window (it fits the device), is pruned to the budget of its own row count,
    merged into the resident summary, and the resident summary is pruned to the budget of the rows seen so far."""
    sketch, accumulated = None, 0
    for x in batches:
        x = np.asarray(x, dtype=np.float32)
        if sketch is None:
            sketch = SketchContainer(x.shape[1], max_bin)
        accumulated += x.shape[0]
        sketch.push(_sorted_columns(x), required_cuts_per_column(max_bin, x.shape[0]), x.shape[0])
        sketch.prune(required_cuts_per_column(max_bin, accumulated))
    return sketch.make_cuts(), sketch
