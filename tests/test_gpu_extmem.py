"""External memory: the quantised matrix in pinned HOST memory, streamed page by page through
two device staging buffers by every phase of the updater that reads it (b200hist_ellpack_create_paged).  The bar: the
trees, positions and prediction cache of the in-core matrix, bit for bit -- integer histograms do not depend on how rows
are grouped into pages.  Reference: src/tree/updater_gpu_hist.cu:475-499,529-580, src/data/ellpack_page_source.h:128-337."""
import json
import os

import numpy as np
import pytest

from tests.helpers import make_cuts, make_regression
from tests.test_gpu_parity import xb  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,f,max_bin,missing,page_rows,params", [
    (60000, 40, 64, 0.0, 7001, dict(max_depth=6)),                      # 9 ragged pages, pair kernel after tree 1
    (60000, 40, 64, 0.05, 20000, dict(max_depth=6)),                    # null symbol, 3 even pages
    (30000, 12, 300, 0.1, 4097, dict(max_depth=5)),                     # uint16 symbols
    (30000, 8, 600, 0.0, 9999, dict(max_depth=4)),                      # generic kernel (> 440 symbols)
    (50000, 70, 32, 0.0, 50000, dict(max_depth=7)),                     # ONE page (everything staged once per phase)
    (50000, 33, 32, 0.0, 1000, dict(max_depth=0, max_leaves=24, grow_policy="lossguide")),   # 50 pages, pages < a tile
], ids=["ragged9", "missing3", "u16", "generic", "one_page", "lossguide50"])
def test_paged_matrix_trains_like_the_in_core_matrix(xb, n, f, max_bin, missing, page_rows, params):
    import torch
    x, y = make_regression(n, f, seed=131, missing=missing)
    cp, cv = make_cuts(x, max_bin)
    cuts = xb.HistogramCuts(cp, cv)
    xd = torch.from_numpy(x).cuda()
    yd = torch.from_numpy(y).cuda()
    core = xb.EllpackPage.from_dense(xd, cuts)
    paged = xb.EllpackPage.from_dense(xd, cuts, page_rows=page_rows)
    info = paged.paging_info()
    assert info["n_pages"] == -(-n // page_rows) and core.paging_info()["n_pages"] == 0
    ups = [xb.GPUHistMaker(core), xb.GPUHistMaker(paged)]
    for u in ups:
        u.configure(dict(params, eta=0.4))
    margins = [torch.zeros(n, device="cuda") for _ in ups]
    poss = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in ups]
    for it in range(3):
        trees = []
        for u, m, p in zip(ups, margins, poss):
            if it == 2:   # hessians with low words: the general accumulation path
                pr = torch.sigmoid(m)
                gp = torch.stack([pr - (yd > 0).float(), pr * (1 - pr)], dim=1).contiguous()
            else:
                gp = torch.stack([m - yd, torch.ones_like(yd)], dim=1).contiguous()
            trees.append(u.update(gp, p))
            u.update_prediction_cache(m)
        assert trees[0].num_nodes == trees[1].num_nodes and trees[0].num_nodes > 3
        assert trees[0].array.tobytes() == trees[1].array.tobytes(), "tree %d differs" % it
        assert torch.equal(poss[0], poss[1]), "positions of tree %d differ" % it
        assert torch.equal(margins[0], margins[1])
        for nidx in (0, 1, 2):
            assert np.array_equal(ups[0].node_histogram(nidx), ups[1].node_histogram(nidx))
    assert paged.paging_info()["h2d_bytes"] > 0


def test_external_memory_dmatrix_beyond_the_resident_budget_through_the_shim(cuda_device, monkeypatch):
    """ExtMemQuantileDMatrix of the unmodified reference (five pages in its host cache) with a resident budget of 1 MB for
    our side: the shim builds a paged page set (pinned host image, ~5 000-row staging buffers) and the reference Booster
    must end on the model its own updater builds from the same DMatrix."""
    import ctypes
    import torch
    from oracle import ref_capi as rc
    from tests import test_shim_dropin as sd
    if not (rc.available("cuda") and os.path.exists(sd.SHIM)):
        pytest.skip("needs oracle/_ref_cuda/libxgboost.so and the shim")
    sd._ensure_shim(rc)
    import xgboost_b200 as xbm
    x, y = make_regression(70000, 48, seed=133, missing=0.04)
    parts = [x[:15000], x[15000:15003], x[15003:40000], x[40000:61111], x[61111:]]

    def train(updater, budget):
        if budget:
            monkeypatch.setenv("B200HIST_RESIDENT_BUDGET_BYTES", str(budget))
        else:
            monkeypatch.delenv("B200HIST_RESIDENT_BUDGET_BYTES", raising=False)
        dm = rc.DMatrix.extmem_quantile_from_batches([torch.from_numpy(p).cuda() for p in parts], max_bin=64, kind="cuda")
        dm.set_label(y)
        bst = rc.Booster({"tree_method": "hist", "device": "cuda:0", "max_bin": 64, "max_depth": 7, "eta": 0.3,
                          "objective": "reg:squarederror", "base_score": 0.5, "updater": updater}, dm)
        for i in range(4):
            bst.update(i)
        trees = bst.save_json()["learner"]["gradient_booster"]["model"]["trees"]
        keep = ("left_children", "right_children", "split_indices", "split_conditions", "default_left",
                "base_weights", "loss_changes", "sum_hessian")
        return [{k: t[k] for k in keep} for t in trees], bst.predict_margin(dm)

    ref_trees, ref_margin = train("grow_gpu_hist_reference", 0)
    before = xbm.launch_count()
    our_trees, our_margin = train("grow_b200_hist", 1 << 20)          # 1 MB: 70000 x 64 B = 4.5 MB does not fit
    launches_paged = xbm.launch_count() - before
    before = xbm.launch_count()
    res_trees, res_margin = train("grow_b200_hist", 1 << 30)          # fits: one resident page, far fewer launches
    launches_resident = xbm.launch_count() - before
    assert json.dumps(our_trees) == json.dumps(ref_trees)
    assert np.array_equal(our_margin.view(np.uint32), ref_margin.view(np.uint32))
    assert json.dumps(res_trees) == json.dumps(ref_trees)
    assert launches_paged > 2 * launches_resident, (launches_paged, launches_resident)
