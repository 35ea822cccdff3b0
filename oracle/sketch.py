"""CPU restatement of the reference GPU quantile sketch -> HistogramCuts, for the in-core, unweighted,
single-window case (SimpleDMatrix on the device, all entries fit one sketch batch).

TEST INFRASTRUCTURE ONLY (see oracle/hist_oracle.c header).  Follows, relative to /root/reference:
  DeviceSketchWithHessian / ProcessWeightedBatch     src/common/hist_util.cu:234-397
  RequiredSampleCutsPerColumn                        src/common/hist_util.cu:31-35
  SketchEpsilon, WQuantileSketch::LimitSizeLevel     src/common/quantile.h:579-601,734-738  (kFactor = 8)
  SketchContainer::Push -> PruneImpl                 src/common/quantile.cu:108-153,340-378
  BinarySearchQueryIndex                             src/common/quantile.cu:36-61
  SketchContainer::ScanInput (dedupe)                src/common/quantile.cu:380-418
  SketchContainer::MakeCuts -> WQSummary::QueryCutValues   src/common/quantile.cu:658-709, quantile.h:280-340
Pinned by the cut values the reference itself produced on a B200 (tests/golden/ref_gpu_trees.json).
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
