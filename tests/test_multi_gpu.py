"""Multi-GPU paths (feature-block sharded histogram + one exchange per level; row split + one summing exchange per level).

CPU part (no GPU): the host logic of the sharding under a real world_size=2 torch.distributed (gloo)
job -- group ownership must tile the group range exactly and the unique-id hand-off must deliver the same
128 bytes to every rank.
GPU part (needs >= 2 GPUs, run with `gpurun --gpus 2`): a 2-rank NCCL job must produce trees that are
bit-identical to the single-GPU trees.
"""
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT

_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
dist.init_process_group("gloo")
import xgboost_b200 as xb
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
for n_groups in (1, 2, 3, 7, 8, 11, 64):
    b, e = xb.shard_groups(n_groups, world, rank)
    t = torch.tensor([b, e], dtype=torch.int64)
    out = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(out, t)
    spans = [tuple(o.tolist()) for o in out]
    # contiguous, ordered, exact cover, balanced within one group
    ok &= spans[0][0] == 0 and spans[-1][1] == n_groups
    ok &= all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    sizes = [s[1] - s[0] for s in spans]
    ok &= max(sizes) - min(sizes) <= 1
# control-plane hand-off of the 128-byte id (content is arbitrary here: NCCL itself needs a GPU)
ident = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
dist.broadcast(ident, src=0)
ok &= ident.tolist() == list(range(128))
flag = torch.tensor([1 if ok else 0])
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
'''


def test_feature_shard_host_logic_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_shard_groups_single_process():
    import xgboost_b200 as xb
    assert xb.shard_groups(8, 1, 0) == (0, 8)
    assert [xb.shard_groups(8, 8, r) for r in range(8)] == [(r, r + 1) for r in range(8)]
    assert [xb.shard_groups(3, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 3)]
    with pytest.raises(xb.B200HistError):
        xb.shard_groups(8, 2, 2)


_GPU_WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import xgboost_b200 as xb
from tests.helpers import make_cuts, make_regression, squared_error_gpair
x, y = make_regression(60000, 96, seed=17)
cp, cv = make_cuts(x, 256)
page = xb.EllpackPage.from_dense(torch.from_numpy(x).to(dev), xb.HistogramCuts(cp, cv))
gp = torch.from_numpy(squared_error_gpair(np.zeros_like(y), y)).to(dev)
up = xb.GPUHistMaker(page); up.configure(dict(max_depth=6))
single = up.update(gp)
h_single = [up.node_histogram(i) for i in (0, 1, 2)]
comm = xb.Communicator.from_torch_distributed(dev)
up.set_comm(comm)
assert up.num_feature_groups() == 3
sharded = up.update(gp)
ok = single.structure() == sharded.structure()
ok &= all(np.array_equal(up.node_histogram(i), h) for i, h in zip((0, 1, 2), h_single))
ok &= [n["leaf_value"] for n in single.nodes] == [n["leaf_value"] for n in sharded.nodes]
# the default collective: peer-memory push over NVLink P2P (a second updater; the first keeps NCCL)
up2 = xb.GPUHistMaker(page); up2.configure(dict(max_depth=6)); up2.enable_peer_exchange()
for _ in range(3):  # repeated trees exercise the start-of-tree barrier
    pushed = up2.update(gp)
    ok &= single.structure() == pushed.structure()
    ok &= all(np.array_equal(up2.node_histogram(i), h) for i, h in zip((0, 1, 2), h_single))
# row split: every rank holds ITS rows (uneven slices), full-width partial histograms summed over peer memory
world = dist.get_world_size()
cutpoints = [0] + [int(60000 * (0.4 + 0.6 * (r + 1) / world)) if r + 1 < world else 60000 for r in range(world)]
lo, hi = cutpoints[rank], cutpoints[rank + 1]
page_r = xb.EllpackPage.from_dense(torch.from_numpy(x[lo:hi]).to(dev), xb.HistogramCuts(cp, cv))
up3 = xb.GPUHistMaker(page_r); up3.configure(dict(max_depth=6)); up3.enable_peer_exchange(row_split=True)
pos_single = torch.empty(60000, dtype=torch.int32, device=dev)
single_again = up.update(gp, pos_single)  # NCCL feature-shard updater: same tree, full position vector
for it in range(3):  # trees 2 and 3 go through the 64-feature pair kernel
    pos_r = torch.empty(hi - lo, dtype=torch.int32, device=dev)
    rs = up3.update(gp[lo:hi].contiguous(), pos_r)
    ok &= single.structure() == rs.structure()
    ok &= [n["leaf_value"] for n in single.nodes] == [n["leaf_value"] for n in rs.nodes]
    ok &= [n["sum_hess"] for n in single.nodes] == [n["sum_hess"] for n in rs.nodes]
    ok &= all(np.array_equal(up3.node_histogram(i), h) for i, h in zip((0, 1, 2), h_single))
    ok &= bool(torch.equal(pos_r, pos_single[lo:hi]))
# logistic gradients: hessians with low words (general accumulation path), another tree shape
p0 = torch.sigmoid(torch.from_numpy(y).to(dev) * 0.3)
gl = torch.stack([p0 - (torch.from_numpy(y).to(dev) > 0).float(), p0 * (1 - p0)], dim=1).contiguous()
ref_l = up.update(gl)
rs_l = up3.update(gl[lo:hi].contiguous())
ok &= ref_l.structure() == rs_l.structure()
ok &= [n["leaf_value"] for n in ref_l.nodes] == [n["leaf_value"] for n in rs_l.nodes]
# teardown in two steps: unmap everywhere, barrier (inside), then the updaters may go
up3.disable_peer_exchange()
up2.disable_peer_exchange()
t_after = up2.update(gp)   # back to single-GPU operation: the plain tree again
ok &= single.structure() == t_after.structure()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
'''


@pytest.mark.gpu
def test_two_rank_nccl_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    script = tmp_path / "gpu_worker.py"
    script.write_text(_GPU_WORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29614", str(script)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [{}, {"objective": "binary:logistic", "max_depth": 5},
                                   {"subsample": 0.7, "colsample_bytree": 0.6, "grow_policy": "lossguide",
                                    "max_leaves": 20, "max_depth": 0}],
                         ids=["squarederror", "logistic", "sampled_lossguide"])
def test_two_process_reference_job_through_the_shim(tmp_path, extra):
    """Distributed mode through the boundary: a 2-worker job of the UNMODIFIED reference (its own RabitTracker and
    communicator, row-split data, one GPU per worker, uneven shards) with our factory serving grow_gpu_hist must give
    the model the reference's own distributed gpu_hist gives (histogram.cu:598-609 AllReduceHist, quantiser.cu:88-123
    GlobalSum, updater_gpu_hist.cu:661-668 root sum) -- same trees bit for bit on both workers."""
    import json
    import torch
    from oracle import ref_capi as rc
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    shim = os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist_xgb_shim.so")
    if not (rc.available("cuda") and os.path.exists(shim)):
        pytest.skip("needs oracle/_ref_cuda/libxgboost.so and the shim")
    models = {}
    for updater in ("grow_gpu_hist_reference", "grow_b200_hist"):
        tracker = rc.Tracker(2, kind="cuda")
        args = tracker.worker_args()
        out = str(tmp_path / ("model_%s.json" % updater))
        procs = [subprocess.Popen([sys.executable, "-m", "tests.dist_shim_worker", str(t), json.dumps(args), updater, out,
                                   json.dumps(extra)], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                  text=True) for t in range(2)]
        logs = [p.communicate(timeout=600)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(l[-3000:] for l in logs)
        tracker.wait(60)
        tracker.free()
        got = [json.load(open("%s.rank%d" % (out, r))) for r in range(2)]
        assert got[0]["trees"] == got[1]["trees"], "the two workers hold different models"
        models[updater] = got
    ref, ours = models["grow_gpu_hist_reference"], models["grow_b200_hist"]
    assert ours[0]["launches"] > 0 and ours[1]["launches"] > 0, "the shim did not route Update() into libb200hist.so"
    assert ref[0]["launches"] == 0
    assert len(ref[0]["trees"]) == 4 and len(ref[0]["trees"][0]["left_children"]) > 7
    assert json.dumps(ours[0]["trees"]) == json.dumps(ref[0]["trees"])
    for r in range(2):
        assert ours[r]["margin_bits"] == ref[r]["margin_bits"], "prediction cache of worker %d" % r
