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
