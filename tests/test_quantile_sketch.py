"""Q2, batch by batch: the resident quantile sketch behind QuantileDMatrix(DataIter).

tests/golden/ref_quantile_dmatrix_cuts.json holds the cut values the UNMODIFIED reference produced for
`XGQuantileDMatrixCreateFromCallback` over streams of device batches (tools/make_golden_gpu.py --sketch-out on a
B200): several equal batches, uneven batches with missing values / a discrete and a constant column, a single
column (where Prune(IntermediateNumCuts) bites), batches and columns that are entirely missing.

 * CPU test: oracle/sketch.py (Push / ScanInput / Merge / FixError / Prune / MakeCuts restated in numpy float32)
   reproduces the reference's cuts bit for bit -- this pins the restatement.
 * GPU tests: b200hist_sketch_{create,push_dense,make_cuts} reproduce the reference's cuts bit for bit, agree
   with the oracle on further streams, and keep the properties the reference's own tests check
   (tests/cpp/common/test_hist_util.h:134-180: strictly increasing cuts, balanced bins).
"""
import json
import os

import numpy as np
import pytest

from tests.helpers import SKETCH_CASES, make_sketch_batches

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_quantile_dmatrix_cuts.json")


def _golden():
    cases = {c["name"]: c for c in json.load(open(GOLD))["cases"]}
    return [(c, cases[c["name"]]) for c in SKETCH_CASES]


def _cuts_of(gold):
    return (np.array(gold["cut_ptrs"], dtype=np.uint32),
            np.array(gold["cut_values_bits"], dtype=np.uint32).view(np.float32))


def _assert_same_cuts(ptrs, vals, cp, cv):
    assert np.array_equal(np.asarray(ptrs, dtype=np.uint32), cp)
    vals = np.asarray(vals, dtype=np.float32)
    # -0.0 / +0.0 are equal keys for the reference's comparison sort: the sign of a zero cut is unspecified
    assert np.array_equal(vals, cv)
    nz = cv != 0
    assert np.array_equal(vals[nz].view(np.uint32), cv[nz].view(np.uint32))


def test_golden_covers_every_case():
    names = {c["name"] for c in json.load(open(GOLD))["cases"]}
    assert names == {c["name"] for c in SKETCH_CASES}


@pytest.mark.parametrize("case,gold", _golden() if os.path.exists(GOLD) else [], ids=lambda v: v.get("name", ""))
def test_oracle_reproduces_reference_quantile_dmatrix_cuts(case, gold):
    from oracle import sketch
    (p, v), cont = sketch.quantile_dmatrix_cuts(make_sketch_batches(case), case["max_bin"])
    assert cont.fix_error_changes == 0   # FixError stayed a no-op: the result does not depend on thread order
    _assert_same_cuts(p, v, *_cuts_of(gold))


def test_oracle_single_window_is_one_push():
    """The in-core DeviceSketch (one window) and a one-batch QuantileDMatrix give the same cuts."""
    from oracle import sketch
    rng = np.random.default_rng(5)
    x = rng.standard_normal((20000, 4)).astype(np.float32)
    x[:, 1] = np.round(x[:, 1] * 3)
    p0, v0 = sketch.device_sketch_cuts(x, 64)
    (p1, v1), _ = sketch.quantile_dmatrix_cuts([x], 64)
    assert np.array_equal(p0, p1) and np.array_equal(v0.view(np.uint32), v1.view(np.uint32))


def test_oracle_merge_of_disjoint_and_equal_summaries():
    """MergeImpl's three cases on literal summaries: x before y, equal values combine, y before x
    (formulas of quantile.cu:268-322)."""
    from oracle import sketch
    f = np.float32
    x = np.array([[0, 1, 1, 1.0], [1, 2, 1, 3.0], [2, 4, 2, 5.0]], dtype=f)   # 4 rows: 1, 3, 5, 5
    y = np.array([[0, 2, 2, 3.0], [2, 3, 1, 4.0]], dtype=f)                   # 3 rows: 3, 3, 4
    m = sketch._unique_by_value(sketch.merge_columns(x, y))
    assert np.array_equal(m[:, 3], np.array([1, 3, 4, 5], dtype=f))
    # value 3 is in both: ranks and weights add; value 4 (from y) sits after x's {1, 3} and before x's 5
    assert np.array_equal(m[1], np.array([1 + 0, 2 + 2, 1 + 2, 3.0], dtype=f))
    assert np.array_equal(m[2], np.array([2 + 2, 3 + (4 - 2), 1, 4.0], dtype=f))
    assert np.array_equal(m[3], np.array([2 + 3, 4 + 3, 2, 5.0], dtype=f))
    assert m[-1, 1] == 7                                                        # total weight preserved


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
def _cuda_cuts(batches, max_bin):
    import torch
    import xgboost_b200 as xb
    sk = xb.QuantileSketch(batches[0].shape[1], max_bin, torch.device("cuda", 0))
    for b in batches:
        sk.push(torch.from_numpy(b).cuda())
    cuts = sk.make_cuts()
    fixes = sk.fix_error_changes
    sk.close()
    return cuts, fixes


@pytest.mark.gpu
@pytest.mark.parametrize("case,gold", _golden() if os.path.exists(GOLD) else [], ids=lambda v: v.get("name", ""))
def test_cuda_reproduces_reference_quantile_dmatrix_cuts(case, gold, cuda_device):
    cuts, fixes = _cuda_cuts(make_sketch_batches(case), case["max_bin"])
    assert fixes == 0
    _assert_same_cuts(cuts.cut_ptrs, cuts.cut_values, *_cuts_of(gold))


@pytest.mark.gpu
@pytest.mark.parametrize("seed,f,max_bin,rows", [(1, 3, 8, [1, 1, 1, 2]), (2, 7, 32, [257, 4096, 31, 9000]),
                                                (3, 2, 256, [30000, 30000, 30000]), (4, 40, 64, [2500] * 5)])
def test_cuda_stream_matches_oracle(seed, f, max_bin, rows, cuda_device):
    from oracle import sketch
    rng = np.random.default_rng(seed)
    batches = []
    for b, r in enumerate(rows):
        x = rng.standard_normal((r, f)).astype(np.float32) * np.float32(1 + b)
        x[:, 0] = np.round(x[:, 0])
        if f > 2:
            x[rng.random(x.shape) < 0.05] = np.nan
        batches.append(x)
    (p, v), _ = sketch.quantile_dmatrix_cuts(batches, max_bin)
    cuts, _ = _cuda_cuts(batches, max_bin)
    _assert_same_cuts(cuts.cut_ptrs, cuts.cut_values, p, v)


@pytest.mark.gpu
def test_cuda_stream_properties_at_scale(cuda_device):
    """40 batches x 250k rows (10M rows per column, past what the numpy oracle finishes quickly): cuts strictly
    increase and every bin holds about N / max_bin rows (sketch error budget eps = 1 / max_bin)."""
    import torch
    import xgboost_b200 as xb
    dev = torch.device("cuda", 0)
    f, max_bin, nb, rows = 16, 256, 40, 250_000
    sk = xb.QuantileSketch(f, max_bin, dev)
    gen = torch.Generator(device=dev).manual_seed(99)
    counts = None
    batches = []
    for b in range(nb):
        x = torch.randn((rows, f), generator=gen, device=dev)
        x[:, 1] = torch.exp(x[:, 1])       # skewed column
        sk.push(x)
        batches.append(x[:, :2].clone())
    cuts = sk.make_cuts()
    assert sk.fix_error_changes == 0
    assert int(cuts.cut_ptrs[-1]) == f * max_bin
    for j in range(f):
        seg = cuts.cut_values[cuts.cut_ptrs[j]:cuts.cut_ptrs[j + 1]]
        assert np.all(np.diff(seg) > 0)
    total = nb * rows
    for j in range(2):
        seg = torch.from_numpy(cuts.cut_values[cuts.cut_ptrs[j]:cuts.cut_ptrs[j + 1]].copy()).to(dev)
        col = torch.cat([b[:, j] for b in batches])
        bins = torch.bucketize(col, seg, right=True).clamp_(max=max_bin - 1)
        counts = torch.bincount(bins, minlength=max_bin).cpu().numpy()
        ideal = total / max_bin
        assert counts.max() <= 1.5 * ideal and counts.min() >= 0.5 * ideal, (j, counts.min(), counts.max(), ideal)


@pytest.mark.gpu
def test_cuda_sketch_rejects_oversized_window(cuda_device):
    import torch
    import xgboost_b200 as xb
    sk = xb.QuantileSketch(2, 16, torch.device("cuda", 0))
    with pytest.raises(ValueError):
        sk.push(torch.zeros((4, 3), device="cuda"))
    sk.push(torch.zeros((0, 2), device="cuda"))   # an empty batch is a no-op, like an empty DataIter batch
    sk.push(torch.ones((5, 2), device="cuda"))
    cuts = sk.make_cuts()
    assert list(cuts.cut_ptrs) == [0, 1, 2]
