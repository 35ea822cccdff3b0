"""Shared helpers for the parity tests: seeded synthetic data, simple exact-quantile cuts
(test plumbing, not the product's sketch) and tree comparison."""
import numpy as np


def make_cuts(x: np.ndarray, max_bin: int):
    """Per-feature cuts from exact quantiles of x (NaN ignored); last cut is a sentinel above the max,
    like the reference's `last + (|last| + 1e-5)` (src/common/quantile.h:280-340)."""
    n, f = x.shape
    ptrs = [0]
    vals = []
    for j in range(f):
        col = x[:, j]
        col = col[~np.isnan(col)]
        u = np.unique(col.astype(np.float32))
        if u.size == 0:
            cuts = np.array([1e-5], dtype=np.float32)
        else:
            if u.size > max_bin - 1:
                qs = np.quantile(col, np.linspace(0, 1, max_bin + 1)[1:-1]).astype(np.float32)
                c = np.unique(qs)
            else:
                c = u[1:] if u.size > 1 else np.array([], dtype=np.float32)
            last = np.float32(u[-1])
            sentinel = np.float32(last + (abs(last) + np.float32(1e-5)))
            cuts = np.concatenate([c[c < sentinel], [sentinel]]).astype(np.float32)
        vals.append(cuts)
        ptrs.append(ptrs[-1] + cuts.size)
    return np.array(ptrs, dtype=np.uint32), np.concatenate(vals).astype(np.float32)


def make_regression(n, f, seed, missing=0.0, dtype=np.float32):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, f)).astype(dtype)
    w = rng.standard_normal(f).astype(dtype)
    # NOT `x @ w`: BLAS float32 matmul rounds differently on different CPUs, and the golden fixtures must
    # regenerate bit-identically everywhere.  Element-wise IEEE ops in a fixed order are reproducible.
    acc = np.zeros(n, dtype=np.float64)
    for j in range(f):
        acc += x[:, j].astype(np.float64) * np.float64(w[j])
    y = (acc + 0.1 * rng.standard_normal(n)).astype(dtype)
    if missing > 0:
        mask = rng.random((n, f)) < missing
        x = x.copy()
        x[mask] = np.nan
    return x, y


def make_sparse_regression(n, f, density, seed):
    """Deterministic sparse problem: (dense-with-NaN matrix, y, CSR indptr/indices/values)."""
    rng = np.random.default_rng(seed)
    mask = rng.random((n, f)) < density
    vals_full = rng.standard_normal((n, f)).astype(np.float32)
    x = np.where(mask, vals_full, np.float32(np.nan)).astype(np.float32)
    w = rng.standard_normal(f).astype(np.float32)
    acc = np.zeros(n, dtype=np.float64)
    for j in range(f):
        acc += np.where(mask[:, j], vals_full[:, j], np.float32(0)).astype(np.float64) * np.float64(w[j])
    y = (acc + 0.1 * rng.standard_normal(n)).astype(np.float32)
    indptr = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int64)
    indices = np.nonzero(mask)[1].astype(np.int32)
    values = x[mask].astype(np.float32)
    return x, y, indptr, indices, values


def squared_error_gpair(pred, y):
    g = (pred - y).astype(np.float32)
    h = np.ones_like(g)
    return np.stack([g, h], axis=1).astype(np.float32)


def logistic_gpair(margin, y):
    p = 1.0 / (1.0 + np.exp(-margin.astype(np.float64)))
    g = (p - y).astype(np.float32)
    h = np.maximum(p * (1 - p), 1e-16).astype(np.float32)
    return np.stack([g, h], axis=1).astype(np.float32)


def f32_bits(v):
    return int(np.float32(v).view(np.uint32))


# Batch streams for the QuantileDMatrix sketch (multi-window) parity cases; regenerated from the seed everywhere.
SKETCH_CASES = [
    dict(name="qdm_8x5000_b16", seed=201, f=6, max_bin=16, batch_rows=[5000] * 8),
    dict(name="qdm_uneven_b64", seed=202, f=5, max_bin=64, batch_rows=[1000, 20000, 7, 15000, 13993], missing=0.1,
         discrete_col=2, const_col=4),
    dict(name="qdm_10x20000_b256", seed=203, f=4, max_bin=256, batch_rows=[20000] * 10),
    dict(name="qdm_single_col_b32", seed=204, f=1, max_bin=32, batch_rows=[3000] * 12),
    dict(name="qdm_holes_b32", seed=205, f=3, max_bin=32, batch_rows=[4000, 500, 4000, 6000], nan_batches=[1],
         nan_cols={0: [1], 2: [1, 2]}),
    dict(name="qdm_6x50000_b256_wide", seed=206, f=12, max_bin=256, batch_rows=[50000] * 6, discrete_col=3),
]


def make_sketch_batches(case):
    """List of float32 matrices (NaN = missing) for a SKETCH_CASES entry."""
    rng = np.random.default_rng(case["seed"])
    out = []
    for b, rows in enumerate(case["batch_rows"]):
        x = rng.standard_normal((rows, case["f"]), dtype=np.float32)
        x = (x * np.float32(1.0 + 0.25 * b) + np.float32(0.1 * b)).astype(np.float32)   # drifting batches
        if case.get("discrete_col") is not None:
            x[:, case["discrete_col"]] = np.round(x[:, case["discrete_col"]] * 4.0)
        if case.get("const_col") is not None:
            x[:, case["const_col"]] = np.float32(2.5)
        if case.get("missing"):
            x[rng.random(x.shape) < case["missing"]] = np.nan
        if b in case.get("nan_batches", []):
            x[:] = np.nan
        for col in case.get("nan_cols", {}).get(b, []):
            x[:, col] = np.nan
        out.append(np.ascontiguousarray(x))
    return out


class MT19937:
    """std::mt19937 (the reference's RandomEngine, include/xgboost/context.h) seeded like `engine.seed(s)`;
    next() returns the successive 32-bit outputs -- e.g. the per-tree seeds the GPU row sampler draws with ctx->Rng()()."""

    def __init__(self, seed: int):
        self.mt = [0] * 624
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.idx = 624

    def next(self) -> int:
        if self.idx >= 624:
            mt = self.mt
            for k in range(624):
                y = (mt[k] & 0x80000000) | (mt[(k + 1) % 624] & 0x7FFFFFFF)
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.idx = 0
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def copy(self) -> "MT19937":
        c = MT19937.__new__(MT19937)
        c.mt, c.idx = list(self.mt), self.idx
        return c


# ---------------------------------------------------------------------------------------------------------------------
# Data recipes of the fixtures generated by tools/probe_interaction_reference.py and tools/make_golden_cpu_categorical.py
# (kept here so that the generators, the tests and tools/quick_check regenerate exactly the same inputs from a seed)
# ---------------------------------------------------------------------------------------------------------------------
def make_interaction_data(case):
    """n x f standard normal features, y = sum of neighbour products + 0.3 * row sum: a chain of pairwise interactions."""
    rng = np.random.default_rng(case["seed"])
    x = rng.normal(size=(case["n"], case["f"])).astype(np.float32)
    y = (x[:, 0] * x[:, 1] + x[:, 1] * x[:, 2] + x[:, 2] * x[:, 3] + x[:, 3] * x[:, 4] + x[:, 4] * x[:, 5] +
         0.3 * x.sum(1)).astype(np.float32)
    return x, y


CATEGORICAL_N_CATS = [3, 6, 12, 40]


def make_categorical_data(case):
    """Four categorical columns (3 / 6 / 12 / 40 categories, 5 % missing) + three numerical ones; the label is the sum
    of per-category effects, the first numerical column and noise.  Returns x, y, feature types ("c" / "q")."""
    rng = np.random.default_rng(case["seed"])
    n = case["n"]
    cols = [rng.integers(0, k, size=n).astype(np.float32) for k in CATEGORICAL_N_CATS]
    for c in cols:
        c[rng.random(n) < 0.05] = np.nan
    xn = rng.normal(size=(n, 3)).astype(np.float32)
    x = np.column_stack(cols + [xn[:, 0], xn[:, 1], xn[:, 2]]).astype(np.float32)
    eff = [rng.normal(size=k) for k in CATEGORICAL_N_CATS]
    y = sum(e[np.nan_to_num(c, nan=0).astype(int)] for e, c in zip(eff, cols)) + xn[:, 0] + 0.2 * rng.normal(size=n)
    return x, y.astype(np.float32), ["c"] * len(CATEGORICAL_N_CATS) + ["q"] * 3


def pack_reference_page(gbins, cut_ptrs, dense_layout: bool):
    """A matrix of global bins (-1 = missing) as the reference's EllpackPageImpl::gidx_buffer: 8 bytes of padding, then
    an MSB-first bit stream of ceil(log2(n_symbols))-bit symbols (src/common/compressed_iterator.h:37-40,140-152,
    CalcNumSymbols src/data/ellpack_page.cu:140-183).  Dense layout: feature-local bins, null = n_symbols - 1 when any
    value is missing.  Sparse layout: the row's global bins in ascending order, padded with null = total_bins to the
    longest row.  Returns (bytes as uint8 array, row_stride, n_symbols, is_dense)."""
    gb = np.asarray(gbins)
    cp = np.asarray(cut_ptrs, dtype=np.int64)
    n, f = gb.shape
    has_missing = bool((gb < 0).any())
    if dense_layout:
        max_bins = int(np.diff(cp).max())
        n_symbols = max_bins + (1 if has_missing else 0)
        sym = np.where(gb >= 0, gb - cp[:-1][None, :], n_symbols - 1).astype(np.uint64)
        row_stride = f
    else:
        total = int(cp[-1])
        n_symbols = total + 1
        row_stride = max(1, int((gb >= 0).sum(axis=1).max()))
        sym = np.full((n, row_stride), total, dtype=np.uint64)
        for r in range(n):
            present = np.sort(gb[r][gb[r] >= 0])
            sym[r, :present.size] = present
    bits = 1
    while (1 << bits) < n_symbols:
        bits += 1
    flat = sym.reshape(-1)
    stream = ((flat[:, None] >> np.arange(bits - 1, -1, -1, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint8)
    packed = np.packbits(stream.reshape(-1))
    buf = np.concatenate([np.zeros(8, dtype=np.uint8), packed, np.zeros(8, dtype=np.uint8)])
    return buf, row_stride, n_symbols, not has_missing
