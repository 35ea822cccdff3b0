#!/usr/bin/env python
"""bench.py -- boosting iterations/second of the B200 hist tree updater on the reference's headline
workload (BASELINE.json: 100M x 256 dense f32, 256 bins, depth 8), synthetic data.

  python bench.py --gpus N --steps K --warmup W            our arm
  python bench.py --impl reference --gpus N ...            the reference's CPU `hist` path on the host cores

One "step" = one boosting iteration on the resident quantised matrix: objective gradient (reg:squarederror)
-> TreeUpdater::Update (quantise, histograms, split evaluation, row partition, finalise) ->
UpdatePredictionCache.  `value` times it with every input resident in HBM; `e2e` times the same iteration
through the C-ABI entry that takes the gradient from (pinned) HOST memory and returns the tree to the host.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "boosting iters/sec, 100Mx256 dense f32, 256 bins, depth 8"
UNIT = "iters/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=int(os.environ.get("B200_BENCH_ROWS", 100_000_000)))
    ap.add_argument("--features", type=int, default=int(os.environ.get("B200_BENCH_FEATURES", 256)))
    ap.add_argument("--max-bin", type=int, default=256)
    ap.add_argument("--depth", type=int, default=int(os.environ.get("B200_BENCH_DEPTH", 8)))
    ap.add_argument("--cpu-sample-rows", type=int, default=int(os.environ.get("B200_BENCH_CPU_ROWS", 4_000_000)),
                    help="rows of the workload the cpu_baseline leg of OUR arm times (bounded sample)")
    ap.add_argument("--ref-rows", type=int, default=int(os.environ.get("B200_BENCH_REF_ROWS", 25_000_000)),
                    help="rows of the workload `--impl reference` times per step (bounded sample, default a quarter: the "
                         "run ends within minutes; 100000000 = the whole matrix, ~13 min)")
    ap.add_argument("--split", default=os.environ.get("B200_BENCH_SPLIT", "rows"), choices=["rows", "features"],
                    help="multi-GPU decomposition: rows = every rank owns N/P rows, full-width partial histograms summed "
                         "per level over peer memory; features = every rank owns F/P feature blocks of all rows "
                         "(north_star's feature-block shard), slices exchanged per level")
    ap.add_argument("--missing", type=float, default=0.0,
                    help="fraction of entries replaced by NaN (exercises the pages with a null symbol; not the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled every 200 ms during the timed region, through NVML inside this process
    (one library call per sample: no fork of a process that maps 100+ GB, no nvidia-smi start-up in the region).
    Falls back to ONE long-running `nvidia-smi -lms 200` child started before the region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index: int = 0):
        self.index = index
        self.samples = []      # (sm_mhz, sm_max_mhz, set(reasons))
        self._stop = threading.Event()
        self._t = None
        self._child = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._max = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _run(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.samples.append((int(sm), int(self._max), {k for k, b in self.BITS.items() if mask & b}))
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        if self._nvml is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        else:
            self._child = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                            "--format=csv,noheader,nounits", "-lms", "200"],
                                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        if self._child is not None:
            self._child.terminate()
            try:
                out = self._child.communicate(timeout=5)[0]
            except Exception:
                out = ""
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in out.splitlines():
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 6 and parts[0].isdigit() and parts[1].isdigit():
                    self.samples.append((int(parts[0]), int(parts[1]),
                                         {n for n, v in zip(names, parts[2:6]) if v.lower().startswith("active")}))
        sm = sorted(s[0] for s in self.samples)
        mx = [s[1] for s in self.samples]
        reasons = set()
        for s in self.samples:
            reasons |= s[2]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples),
                "source": "nvml" if self._nvml is not None else "nvidia-smi -lms 200"}


# --------------------------------------------------------------------------------------------------
# synthetic workload: X ~ N(0,1) per 1M-row batch generated ON DEVICE, y = X.w + 0.1 eps
# --------------------------------------------------------------------------------------------------
MISSING = 0.0


def batch_of(torch, device, f, rows, batch, b):
    m = min(batch, rows - b * batch)
    g = torch.Generator(device=device).manual_seed(2024 + b)
    x = torch.randn((m, f), generator=g, device=device)
    if MISSING > 0.0:
        x[torch.rand((m, f), generator=g, device=device) < MISSING] = float("nan")
    return x, g


def build_cuts(torch, xb, args, device):
    """Pass 1 of the DMatrix build = what QuantileDMatrix(DataIter of 1M-row device batches) does in the reference
    (quantile_dmatrix.cu:29-126): every batch goes through the resident quantile sketch (b200hist_sketch_push_dense,
    one sketch window per batch); cuts always come from the WHOLE stream.  Returns (cuts, seconds)."""
    n, f = args.rows, args.features
    batch = 1_000_000
    nb = (n + batch - 1) // batch
    t0 = time.perf_counter()
    sketch = xb.QuantileSketch(f, args.max_bin, device)
    for b in range(nb):
        x, _ = batch_of(torch, device, f, n, batch, b)
        sketch.push(x)
        del x
    cuts = sketch.make_cuts()
    sketch.close()
    torch.cuda.synchronize(device)
    return cuts, time.perf_counter() - t0


def build_page(torch, xb, args, device, cuts, batch_lo, batch_hi):
    """Pass 2: regenerate batches [batch_lo, batch_hi) and bin them into an Ellpack page; labels y = X.w + 0.1 eps.
    Returns (page, y, seconds)."""
    n, f = args.rows, args.features
    batch = 1_000_000
    row_lo, row_hi = batch_lo * batch, min(n, batch_hi * batch)
    t0 = time.perf_counter()
    page = xb.EllpackPage(cuts, row_hi - row_lo, device.index, has_missing=MISSING > 0.0)
    gen = torch.Generator(device=device).manual_seed(7)
    w = torch.randn(f, generator=gen, device=device)
    y = torch.empty(row_hi - row_lo, dtype=torch.float32, device=device)
    for b in range(batch_lo, batch_hi):
        x, g = batch_of(torch, device, f, n, batch, b)
        o = b * batch - row_lo
        page.push_dense(x, o)
        y[o:o + x.shape[0]] = torch.nan_to_num(x) @ w + 0.1 * torch.randn(x.shape[0], generator=g, device=device)
        del x
    torch.cuda.synchronize(device)
    return page, y, time.perf_counter() - t0


def algorithmic_bytes(n, f, bin_bytes, depth, row_visits):
    """bytes each stage must move once (dense page)."""
    hist = row_visits * (f * bin_bytes + 16 + 4)
    part = (depth - 1) * n * (4 + bin_bytes + 4)
    quant = n * (8 + 8 + 16)
    final = n * (4 + 4) + n * (4 + 4 + 4)
    return dict(hist=hist, partition=part, quantise=quant, finalise=final, total=hist + part + quant + final)


# --------------------------------------------------------------------------------------------------
# tree signature: every bit the reference records per node (RegTree::ExpandNode + RTreeNodeStat)
# --------------------------------------------------------------------------------------------------
def tree_sha(tree) -> str:
    import hashlib
    a = tree.array
    h = hashlib.sha256()
    for k in ("parent", "left", "right", "findex", "fvalue", "default_left", "is_leaf", "leaf_value", "loss_chg",
              "sum_hess", "base_weight"):
        h.update(a[k].tobytes())
    return h.hexdigest()[:16]


def model_sha(trees_json) -> str:
    import hashlib
    keep = ("left_children", "right_children", "split_indices", "split_conditions", "default_left", "base_weights",
            "loss_changes", "sum_hessian")
    return hashlib.sha256(json.dumps([{k: t[k] for k in keep} for t in trees_json]).encode()).hexdigest()[:16]


# --------------------------------------------------------------------------------------------------
# reference CPU arm (the unmodified reference's tree_method=hist on the host cores)
# --------------------------------------------------------------------------------------------------
def _host_batches(args, rows):
    """The workload's 1M-row float32 batches as numpy arrays held in host RAM (rows x features x 4 bytes; the B200
    hosts have 2 TB), generated ONCE (SURVEY 8(d) generator: torch Philox on the device when there is one -- generation
    is not what is timed -- else numpy on the host) so that both passes of the QuantileDMatrix build read memory."""
    import numpy as np
    f = args.features
    nb = (rows + 999_999) // 1_000_000
    try:
        import torch
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if torch.cuda.is_available() else None
    except Exception:
        torch, dev = None, None
    store = np.empty((rows, f), dtype=np.float32)
    y = np.empty(rows, dtype=np.float32)
    if dev is not None:
        w = torch.randn(f, generator=torch.Generator(device=dev).manual_seed(7), device=dev)
        for b in range(nb):
            x, g = batch_of(torch, dev, f, rows, 1_000_000, b)
            o = b * 1_000_000
            store[o:o + x.shape[0]] = x.cpu().numpy()
            y[o:o + x.shape[0]] = (torch.nan_to_num(x) @ w + 0.1 * torch.randn(x.shape[0], generator=g, device=dev)).cpu().numpy()
            del x
        gen = "torch Philox on cuda, seed 2024 + batch"
    else:
        w = np.random.default_rng(7).standard_normal(f).astype(np.float32)
        for b in range(nb):
            o = b * 1_000_000
            m = min(1_000_000, rows - o)
            rng = np.random.default_rng(2024 + b)
            store[o:o + m] = rng.standard_normal((m, f), dtype=np.float32)
            y[o:o + m] = store[o:o + m] @ w + 0.1 * rng.standard_normal(m, dtype=np.float32)
        gen = "numpy PCG64 on the host, seed 2024 + batch"
    batches = [store[b * 1_000_000:min(rows, (b + 1) * 1_000_000)] for b in range(nb)]
    return batches, y, gen


def reference_cpu(args, steps, warmup, rows):
    """XGBoosterUpdateOneIter of the reference CPU library (oracle/_ref/libxgboost.so, tree_method=hist) on `rows` rows
    of the workload with every host core; per-iteration wall times.  Returns None when the library is not built."""
    import numpy as np
    from oracle import ref_capi as rc
    if not rc.available("cpu"):
        return None
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    batches, y, gen = _host_batches(args, rows)
    dm = rc.DMatrix.quantile_from_batches(batches, max_bin=args.max_bin, kind="cpu", nthread=cores)
    dm.set_label(y)
    build_s = time.perf_counter() - t0
    bst = rc.Booster({"tree_method": "hist", "device": "cpu", "max_depth": args.depth, "max_bin": args.max_bin,
                      "objective": "reg:squarederror", "eta": 0.3, "lambda": 1, "min_child_weight": 1,
                      "nthread": cores}, dm)
    for i in range(warmup):
        bst.update(i)
    ts = []
    t_all = time.perf_counter()
    for i in range(steps):
        t1 = time.perf_counter()
        bst.update(warmup + i)
        ts.append(time.perf_counter() - t1)
    total = time.perf_counter() - t_all
    return dict(rows=rows, cores=cores, s_per_iter_mean=total / max(1, steps), s_per_iter_median=float(np.median(ts)),
                s_per_iter_min=float(min(ts)), s_per_iter_max=float(max(ts)), dmatrix_build_s=build_s, generator=gen,
                omp=os.environ.get("OMP_NUM_THREADS"))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = min(args.rows, args.ref_rows)
    res = reference_cpu(args, args.steps, args.warmup, rows)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libxgboost.so not built"}))
        return
    scale = rows / float(args.rows)                   # 1.0 when the full matrix is timed
    value = scale / res["s_per_iter_mean"]
    sample = ("unmodified reference CPU hist (oracle/_ref/libxgboost.so, XGBoosterUpdateOneIter, nthread=%d, "
              "OMP_NUM_THREADS=%s) on %d x %d rows: mean %.3f s/iter over %d iters (median %.3f, min %.3f, max %.3f)%s"
              % (res["cores"], res["omp"], rows, args.features, res["s_per_iter_mean"], args.steps,
                 res["s_per_iter_median"], res["s_per_iter_min"], res["s_per_iter_max"],
                 "" if scale == 1.0 else
                 "; a step is one iteration on this 1/%g sample, `value` = sample rate x %g -- linear scaling FAVOURS the "
                 "reference: the full 100M x 256 run (profiles/r02_reference_cpu_full_size.json) took 8.14 s/iter on the "
                 "same host = 0.123 iters/s; B200_BENCH_REF_ROWS=100000000 repeats it, ~13 min" % (1.0 / scale, scale)))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup,
            # the wall time of one timed step as it ran (on the sample); 1e3 / value is the full-size equivalent
            "ms_per_step": 1e3 * res["s_per_iter_mean"], "ms_per_step_full_size_equivalent": 1e3 / value,
            "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "synthetic %dx%d dense f32, %d bins, depth %d, reg:squarederror"
                                   % (args.rows, args.features, args.max_bin, args.depth),
                       "timed_rows": rows, "rows_scale": scale, "l2": "inputs exceed L2"},
            "dmatrix_build_s": round(res["dmatrix_build_s"], 2),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# the reference's own Booster on the GPU: e2e of our arm (shim loaded) and the reference gpu_hist beside it
# --------------------------------------------------------------------------------------------------
def reference_booster_section(torch, args, device, steps, warmup, rows=None, features=None, depth=None, missing=None):
    """One QuantileDMatrix of the UNMODIFIED reference CUDA library over the same batches; two Boosters on it:
    updater=grow_b200_hist (our shim: XGBoosterUpdateOneIter = objective gradient + Update + UpdatePredictionCache,
    tree replayed into RegTree on the host) and updater=grow_gpu_hist_reference (the reference's own kernels).
    Wall time per XGBoosterUpdateOneIter call with a device synchronise on both sides."""
    import ctypes
    import numpy as np
    from oracle import ref_capi as rc
    shim = os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist_xgb_shim.so")
    if not (rc.available("cuda") and os.path.exists(shim)):
        return None
    rc.load("cuda")
    ctypes.CDLL(os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist.so"), mode=ctypes.RTLD_GLOBAL)
    keep = ctypes.CDLL(shim, mode=ctypes.RTLD_GLOBAL)
    global MISSING
    n, f = rows or args.rows, features or args.features
    depth = depth or args.depth
    saved_missing = MISSING
    if missing is not None:
        MISSING = missing       # batch_of replaces that fraction of the entries by NaN (a sparse ELLPACK page results)
    nb = (n + 999_999) // 1_000_000
    t0 = time.perf_counter()
    batches = rc.LazyBatches(nb, lambda b: batch_of(torch, device, f, n, 1_000_000, b)[0])
    dm = rc.DMatrix.quantile_from_batches(batches, max_bin=args.max_bin, kind="cuda")
    w = torch.randn(f, generator=torch.Generator(device=device).manual_seed(7), device=device)
    y = np.empty(n, dtype=np.float32)
    for b in range(nb):
        x, g = batch_of(torch, device, f, n, 1_000_000, b)
        yb = torch.nan_to_num(x) @ w + 0.1 * torch.randn(x.shape[0], generator=g, device=device)
        y[b * 1_000_000:b * 1_000_000 + x.shape[0]] = yb.cpu().numpy()
        del x
    dm.set_label(y)
    torch.cuda.synchronize(device)
    build_s = time.perf_counter() - t0
    torch.cuda.empty_cache()
    params = {"tree_method": "hist", "device": "cuda:%d" % device.index, "max_depth": depth,
              "max_bin": args.max_bin, "objective": "reg:squarederror", "eta": 0.3, "lambda": 1, "min_child_weight": 1,
              "base_score": 0.5}
    out = {"dmatrix_build_s": round(build_s, 2)}
    models = {}
    for name, updater in (("ours", "grow_b200_hist"), ("reference_gpu", "grow_gpu_hist_reference")):
        bst = rc.Booster(dict(params, updater=updater), dm)
        ts = []
        for i in range(warmup + steps):
            torch.cuda.synchronize(device)
            t1 = time.perf_counter()
            bst.update(i)
            torch.cuda.synchronize(device)
            ts.append(time.perf_counter() - t1)
        timed = ts[warmup:]
        trees = bst.save_json()["learner"]["gradient_booster"]["model"]["trees"]
        models[name] = model_sha(trees)
        out[name] = {"iters_per_s": len(timed) / sum(timed), "ms_per_iter_mean": 1e3 * sum(timed) / len(timed),
                     "ms_per_iter_all": [round(1e3 * t, 2) for t in ts],
                     "ms_per_iter_median": 1e3 * float(np.median(timed)), "first_iter_ms": 1e3 * ts[0],
                     "steps": steps, "warmup": warmup, "n_trees": len(trees), "model_sha": models[name],
                     "last_tree_nodes": len(trees[-1]["left_children"])}
        del bst
        torch.cuda.empty_cache()
    out["models_identical"] = models["ours"] == models["reference_gpu"]
    MISSING = saved_missing
    del dm, keep
    torch.cuda.empty_cache()
    return out


def reference_booster_distributed(torch, dist, args, device, rank, world, steps, warmup):
    """N > 1: a distributed job of the UNMODIFIED reference -- its own tracker (on rank 0) and communicator, row-split
    data: every rank builds a QuantileDMatrix of ITS batches (the reference's distributed sketch makes the cuts), then
    two Boosters train on it: updater=grow_b200_hist (the shim in distributed mode: IPC handles through the
    reference's collective::Allgather, histograms exchanged over NVLink peer memory) and
    updater=grow_gpu_hist_reference (the reference's own NCCL allreduce).  Wall time per XGBoosterUpdateOneIter,
    max over ranks."""
    import ctypes
    import numpy as np
    from oracle import ref_capi as rc
    shim = os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist_xgb_shim.so")
    if not (rc.available("cuda") and os.path.exists(shim)):
        return None
    rc.load("cuda")
    ctypes.CDLL(os.path.join(ROOT, "xgboost_b200", "lib", "libb200hist.so"), mode=ctypes.RTLD_GLOBAL)
    keep = ctypes.CDLL(shim, mode=ctypes.RTLD_GLOBAL)
    tracker, box = None, [None]
    if rank == 0:
        tracker = rc.Tracker(world, kind="cuda")
        box[0] = tracker.worker_args()
    dist.broadcast_object_list(box, src=0)
    rc.communicator_init("cuda", dict(box[0], dmlc_task_id=str(rank), dmlc_communicator="rabit"))
    out = None
    try:
        n, f = args.rows, args.features
        nb = (n + 999_999) // 1_000_000
        b_lo, b_hi = nb * rank // world, nb * (rank + 1) // world
        t0 = time.perf_counter()
        batches = rc.LazyBatches(b_hi - b_lo, lambda b: batch_of(torch, device, f, n, 1_000_000, b_lo + b)[0])
        dm = rc.DMatrix.quantile_from_batches(batches, max_bin=args.max_bin, kind="cuda")
        w = torch.randn(f, generator=torch.Generator(device=device).manual_seed(7), device=device)
        rows_here = min(n, b_hi * 1_000_000) - b_lo * 1_000_000
        y = np.empty(rows_here, dtype=np.float32)
        for b in range(b_lo, b_hi):
            x, g = batch_of(torch, device, f, n, 1_000_000, b)
            o = (b - b_lo) * 1_000_000
            y[o:o + x.shape[0]] = (torch.nan_to_num(x) @ w + 0.1 * torch.randn(x.shape[0], generator=g, device=device)).cpu().numpy()
            del x
        dm.set_label(y)
        torch.cuda.synchronize(device)
        build_s = time.perf_counter() - t0
        torch.cuda.empty_cache()
        params = {"tree_method": "hist", "device": "cuda:%d" % device.index, "max_depth": args.depth,
                  "max_bin": args.max_bin, "objective": "reg:squarederror", "eta": 0.3, "lambda": 1,
                  "min_child_weight": 1, "base_score": 0.5}
        out = {"dmatrix_build_s": round(build_s, 2)}
        shas = {}
        for name, updater in (("ours", "grow_b200_hist"), ("reference_gpu", "grow_gpu_hist_reference")):
            bst = rc.Booster(dict(params, updater=updater), dm)
            ts = []
            for i in range(warmup + steps):
                torch.cuda.synchronize(device)
                t1 = time.perf_counter()
                bst.update(i)
                torch.cuda.synchronize(device)
                ts.append(time.perf_counter() - t1)
            tot = torch.tensor([sum(ts[warmup:])], device=device, dtype=torch.float64)
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
            trees = bst.save_json()["learner"]["gradient_booster"]["model"]["trees"]
            shas[name] = model_sha(trees)
            out[name] = {"iters_per_s": steps / float(tot.item()), "ms_per_iter_mean": 1e3 * float(tot.item()) / steps,
                         "ms_per_iter_median_rank0": 1e3 * float(np.median(ts[warmup:])),
                         "ms_per_iter_all_rank0": [round(1e3 * t_, 2) for t_ in ts], "steps": steps, "warmup": warmup,
                         "model_sha": shas[name], "last_tree_nodes": len(trees[-1]["left_children"])}
            del bst
            torch.cuda.empty_cache()
        out["models_identical"] = shas["ours"] == shas["reference_gpu"]
        del dm
    finally:
        rc.communicator_finalize("cuda")
        if tracker is not None:
            tracker.wait(120)
            tracker.free()
    del keep
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_b200(args):
    global MISSING
    MISSING = args.missing
    import numpy as np
    import torch
    import torch.distributed as dist

    import xgboost_b200 as xb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    row_split = world > 1 and args.split == "rows"
    n_total, f = args.rows, args.features
    nb = (n_total + 999_999) // 1_000_000
    params = dict(max_depth=args.depth, eta=0.3, reg_lambda=1.0, min_child_weight=1.0)
    cuts, sketch_s = build_cuts(torch, xb, args, device)
    if row_split:  # contiguous runs of 1M-row batches per rank
        page, y, ellpack_s = build_page(torch, xb, args, device, cuts, nb * rank // world, nb * (rank + 1) // world)
    else:
        page, y, ellpack_s = build_page(torch, xb, args, device, cuts, 0, nb)
    build_s = dict(sketch_s=sketch_s, ellpack_s=ellpack_s, sketch_windows=nb)
    info = page.info()
    n = int(info["n_rows"])      # rows resident on THIS rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def make_updater(pg, mode):
        u = xb.GPUHistMaker(pg)
        u.configure(params)
        if mode == "rows":
            u.enable_peer_exchange(row_split=True)
        elif mode == "features":
            if os.environ.get("B200HIST_COMM", "p2p") == "nccl":   # feature blocks + one NCCL allreduce per level
                u._bench_comm = xb.Communicator.from_torch_distributed(device)
                u.set_comm(u._bench_comm)
            else:   # feature blocks + peer-memory push + flag barrier over NVLink (no NCCL in the data path)
                u.enable_peer_exchange()
        return u

    def boost(u, yy, k, keep_pos=None):
        """k boosting iterations from a zero margin; returns the trees."""
        m = yy.shape[0]
        pr = torch.zeros(m, dtype=torch.float32, device=device)
        gp = torch.empty((m, 2), dtype=torch.float32, device=device)
        gp[:, 1] = 1.0
        out = []
        for _ in range(k):
            torch.sub(pr, yy, out=gp[:, 0])
            out.append(u.update(gp, keep_pos))
            u.update_prediction_cache(pr)
        return out

    up = make_updater(page, "single" if world == 1 else ("rows" if row_split else "features"))

    # ---- multi-GPU parity, visible in the bench line: the tree of boosting step VERIFY_STEPS-1 built by all N GPUs
    #      together must hash to the same value as the tree one GPU builds alone from the same cuts and the same rows
    verify = None
    VERIFY_STEPS = 3
    if world > 1:
        trees_n = boost(up, y, VERIFY_STEPS)
        sha_n = tree_sha(trees_n[-1])
        gathered = [None] * world
        dist.all_gather_object(gathered, sha_n)
        sha_1 = None
        if rank == 0:
            if row_split:
                full_page, full_y, _ = build_page(torch, xb, args, device, cuts, 0, nb)
            else:
                full_page, full_y = page, y
            solo = make_updater(full_page, "single")
            sha_1 = tree_sha(boost(solo, full_y, VERIFY_STEPS)[-1])
            del solo
            if row_split:
                del full_page, full_y
            torch.cuda.empty_cache()
        barrier()
        verify = {"step": VERIFY_STEPS - 1, "tree_sha": sha_n, "tree_sha_1gpu": sha_1,
                  "all_ranks_agree": len(set(gathered)) == 1, "nodes": trees_n[-1].num_nodes}
        if rank == 0:
            assert verify["all_ranks_agree"], "ranks built different trees: %s" % gathered
            assert sha_n == sha_1, "the %d-GPU tree differs from the single-GPU tree (%s vs %s)" % (world, sha_n, sha_1)

    preds = torch.zeros(n, dtype=torch.float32, device=device)
    gpair = torch.empty((n, 2), dtype=torch.float32, device=device)
    gpair[:, 1] = 1.0
    pos = torch.empty(n, dtype=torch.int32, device=device)

    in_flight = [0]

    def step_resident():
        """One boosting iteration, everything enqueued on the stream: gradient -> Update -> UpdatePredictionCache.  The
        tree's device -> host copy is part of the step but nobody waits for it here: the tree of the PREVIOUS step is
        collected instead (b200hist_updater_update_tree_async / fetch_tree, two pinned staging slots), so the device
        never idles while the host parses a tree.  Returns the tree collected (None in the very first step)."""
        torch.sub(preds, y, out=gpair[:, 0])  # reg:squarederror gradient; hessian stays 1
        up.update_async(gpair, pos)
        in_flight[0] += 1
        up.update_prediction_cache(preds)
        if in_flight[0] == 2:
            in_flight[0] -= 1
            return up.fetch_tree()
        return None

    def drain():
        out = None
        while in_flight[0] > 0:
            in_flight[0] -= 1
            out = up.fetch_tree()
        return out

    # phase events inside the library stay ON in the timed region (~50 cudaEventRecord per tree, no extra sync: the
    # per-tree read-back synchronises anyway), so that the phases, the row visits and the step time are the SAME trees
    up.set_profiling(True)
    for _ in range(args.warmup):
        step_resident()
    drain()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = xb.launch_count()
    prof, host_gap = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tree = step_resident()
        if tree is not None:
            prof.append(up.last_profile())
    tree = drain()                      # the last tree of the timed region: its read-back is inside the region
    prof.append(up.last_profile())
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = xb.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = 1e3 / ms_per_step

    names = [k for k in prof[0] if k not in ("row_visits", "hist_levels")]
    phase = {k: float(np.mean([p[k] for p in prof])) for k in names}
    row_visits_mean = float(np.mean([p["row_visits"] for p in prof]))
    row_visits_minmax = [int(min(p["row_visits"] for p in prof)), int(max(p["row_visits"] for p in prof))]
    hist_levels = None
    if "hist_levels" in prof[0]:
        hist_levels = [float(v) for v in np.mean(np.array([p["hist_levels"] for p in prof]), axis=0)]
    phase_sum = sum(v for k, v in phase.items() if k != "hist_exchange")
    # outside the updater's phases: the objective's gradient kernel and the prediction cache (measured apart, 3 steps)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    parts = []
    for _ in range(3):
        ev[0].record()
        torch.sub(preds, y, out=gpair[:, 0])
        ev[1].record()
        up.update(gpair, pos)
        ev[2].record()
        up.update_prediction_cache(preds)
        ev[3].record()
        torch.cuda.synchronize(device)
        parts.append([ev[i].elapsed_time(ev[i + 1]) for i in range(3)])
    pm = np.median(np.array(parts), axis=0)
    step_parts = {"gradient": float(pm[0]), "prediction_cache": float(pm[2]), "updater_phase_sum_timed_mean": phase_sum,
                  "unattributed": ms_per_step - phase_sum - float(pm[0]) - float(pm[2]),
                  "note": "phases and ms_per_step are means over the SAME timed trees; unattributed = launch gaps + the "
                          "per-tree host read-back"}
    ab = algorithmic_bytes(n, f, info["bin_bytes"], args.depth, row_visits_mean)
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    hist_launches = args.depth
    feat_div = 1 if (world == 1 or row_split) else world   # row split: this rank's visits cover all F features
    hist_bytes_gpu = row_visits_mean * (f * info["bin_bytes"] / feat_div + 16 + 4)
    hist_ms = phase["hist"] - phase.get("hist_exchange", 0.0)   # the kernel itself, without the cross-GPU exchange
    hist_gbs = hist_bytes_gpu / (hist_ms * 1e-3) / 1e9 if hist_ms > 0 else 0.0
    from xgboost_b200 import components as xc
    smem_peak = xc.smem_atomic_peak(True, device)
    sm_clock_hz = (clocks["sm_mhz"] if clocks and clocks.get("sm_mhz") else 1965) * 1e6
    n_sms = torch.cuda.get_device_properties(device).multi_processor_count
    updates_per_s = row_visits_mean * (f / feat_div) / (hist_ms * 1e-3) if hist_ms > 0 else 0.0
    # DRAM traffic of the dominant kernel: only from an ncu --set full capture of THIS workload (profiles/), else null
    traffic, traffic_src = None, "no ncu --set full capture of this workload under profiles/ (null, not a literal)"
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "r02_hist_kernel_ncu_headline.json")))
        if cap.get("rows") == n_total and cap.get("features") == f and world == 1:
            traffic = float(cap["dram_bytes_per_launch_mean"])
            traffic_src = (cap.get("source", "profiles/r02_hist_kernel_ncu_headline.json") +
                           "; the captured launches move %.2f x their own algorithmic bytes"
                           % float(cap.get("traffic_over_algorithmic_captured", float("nan"))))
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "HistKernelPair", "achieved": hist_gbs, "peak": peak, "unit": "GB/s",
                "frac": hist_gbs / peak, "traffic": traffic, "peak_source": peak_src, "traffic_source": traffic_src,
                "launches_per_step": hist_launches, "avg_launch_ms": hist_ms / hist_launches,
                "algorithmic_bytes_per_launch": hist_bytes_gpu / hist_launches,
                "row_visits_per_step_mean": row_visits_mean, "row_visits_per_step_minmax": row_visits_minmax,
                "per_level_ms": hist_levels,
                # the binding resource is the shared-memory atomic pipe, not HBM: the ceiling is MEASURED on this GPU by
                # b200hist_smem_atomic_peak (the kernel's own accumulation sequence with no global loads)
                "smem_atomic_bound": {"updates_per_clk_per_sm": updates_per_s / sm_clock_hz / n_sms,
                                      "ceiling": smem_peak, "frac": updates_per_s / sm_clock_hz / n_sms / smem_peak,
                                      "ceiling_source": "measured: b200hist_smem_atomic_peak(h_lo_zero=1), clock64 based",
                                      "note": "exact int64 accumulation in shared memory bounds the kernel, not HBM"},
                "whole_step_frac": (ab["total"] - ab["hist"] + hist_bytes_gpu) / (ms_per_step * 1e-3) / 1e9 / peak}

    logistic = None
    up.set_profiling(False)

    # ---- everything below is BESIDE the measurement above (other decompositions, end-to-end legs, baselines).  The
    #      JSON line can be built from here on; a watchdog prints it and ends the process should one of those legs hang
    #      (a collective that never returns must not cost the run its line).
    feature_block = e2e = ref_section = other_configs = cpu_baseline = e2e_host = None
    tree_nodes = tree.num_nodes
    parallelism = "single GPU" if world == 1 else (
        "row split x%d: N/P rows per GPU, full-width partial histograms reduce-scattered over NVLink peer memory per tree "
        "level, evaluation by feature slice, 32-byte per-feature winners exchanged" % world
        if row_split else
        "feature-block shard x%d: all rows on every GPU, F/P feature blocks each, one "
        "peer-memory slice exchange per tree level" % world)
    l2_note = "inputs (%.1f GB bin matrix) exceed the 126 MB L2" % (n * info["row_bytes"] / 1e9)

    def make_line(note=None):
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                "config": {"workload": "synthetic %dx%d dense f32, %d bins, depth %d, reg:squarederror" % (n_total, f, args.max_bin, args.depth),
                           "parallelism": parallelism, "l2": l2_note, "tree_nodes": tree_nodes,
                           "tree_readback": "deferred by one step (two pinned staging slots): every step enqueues its "
                                            "tree's D2H copy, the host collects the previous step's tree",
                           "cuts": "own quantile sketch, %d windows of <= 1M rows merged (QuantileDMatrix semantics)"
                                   % build_s["sketch_windows"]},
                "dmatrix_build_s": {"sketch": round(build_s["sketch_s"], 3), "ellpack": round(build_s["ellpack_s"], 3)},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "phase_ms": phase, "step_parts_ms": step_parts,
                "multi_gpu_parity": verify, "feature_block_shard": feature_block,
                "reference_gpu": None if ref_section is None else dict(
                    ref_section["reference_gpu"], unit=UNIT, value=ref_section["reference_gpu"]["iters_per_s"],
                    what="the UNMODIFIED reference's CUDA gpu_hist (oracle/_ref_cuda, updater=grow_gpu_hist_reference) "
                         "through XGBoosterUpdateOneIter on the same QuantileDMatrix in this process (N > 1: its own "
                         "distributed gpu_hist, NCCL allreduce, in the same job): the kernel to beat",
                    dmatrix_build_s=ref_section["dmatrix_build_s"]),
                "e2e_c_abi_host_gradient": e2e_host,
                "other_configs": other_configs, "headline_matrix_binary_logistic": logistic,
                "algorithmic_gb_per_step": {k: v / 1e9 for k, v in ab.items()}}
        if note:
            line["note"] = note
        return line

    def on_deadline():
        if rank == 0:
            sys.stdout.write(json.dumps(make_line("a side leg did not return within its deadline: the line holds what "
                                                  "had been measured by then")) + "\n")
            sys.stdout.flush()
        os._exit(0)

    watchdog = threading.Timer(float(os.environ.get("B200_BENCH_SIDE_DEADLINE_S", 360 if world == 1 else 240)), on_deadline)
    watchdog.daemon = True
    watchdog.start()

    # N > 1, end to end through the C ABI's host entry on the live row-split updater: every step the gradient is
    # computed on the device, copied to pinned HOST memory (where a custom objective would produce it), handed to
    # b200hist_updater_update_tree_host (H2D of the gradient and D2H of the tree inside), the prediction cache is
    # updated and the metric read back.
    if world > 1 and not args.no_e2e:
        try:
            g_host = torch.empty((n, 2), dtype=torch.float32).pin_memory()
            loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

            def step_e2e():
                torch.sub(preds, y, out=gpair[:, 0])
                g_host.copy_(gpair)
                tree_ = up.update(g_host, pos)
                up.update_prediction_cache(preds)
                loss_host.copy_(((preds - y) ** 2).mean().reshape(1))
                return tree_

            for _ in range(2):
                step_e2e()
            barrier()
            k5 = max(1, min(args.steps, 5))
            t0 = time.perf_counter()
            for _ in range(k5):
                tr = step_e2e()
            barrier()
            tt = torch.tensor([(time.perf_counter() - t0) / k5], device=device)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_host = {"value": 1.0 / float(tt.item()), "unit": UNIT, "steps": k5,
                        "h2d_bytes_per_step": int(n_total) * 8, "d2h_bytes_per_step": int(n_total) * 8 + int(tr.num_nodes) * 64 * world + 4 * world,
                        "path": "b200hist_updater_update_tree_host on every rank (row split): gradient device -> pinned host -> "
                                "device, tree -> host, prediction cache, metric -> host; max over ranks"}
            e2e = dict(e2e_host)
            del g_host, loss_host
        except Exception as exc:   # symmetric on every rank; the watchdog covers the asymmetric case
            e2e_host = {"error": repr(exc)}

    # north_star's decomposition measured next to the default one: all rows on every GPU, F/P feature blocks each,
    # one peer-memory slice exchange per tree level (same matrix, same trees)
    feature_block = None
    if row_split and os.environ.get("B200_BENCH_ALSO_FEATURES", "1") != "0":
        up.disable_peer_exchange()   # unmap on every rank, barrier, only then free
        del up, page, preds, gpair, pos
        torch.cuda.empty_cache()
        page_f, y_f, _ = build_page(torch, xb, args, device, cuts, 0, nb)
        up_f = make_updater(page_f, "features")
        sha_f = tree_sha(boost(up_f, y_f, VERIFY_STEPS)[-1])
        nf_rows = int(page_f.info()["n_rows"])
        preds_f = torch.zeros(nf_rows, dtype=torch.float32, device=device)
        gpair_f = torch.empty((nf_rows, 2), dtype=torch.float32, device=device)
        gpair_f[:, 1] = 1.0

        def step_f():
            torch.sub(preds_f, y_f, out=gpair_f[:, 0])
            up_f.update(gpair_f)
            up_f.update_prediction_cache(preds_f)

        for _ in range(3):
            step_f()
        barrier()
        k3 = max(1, min(args.steps, 5))
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(k3):
            step_f()
        f1.record()
        barrier()
        tf = torch.tensor([f0.elapsed_time(f1) / k3], device=device)
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        feature_block = {"value": 1e3 / float(tf.item()), "unit": UNIT, "ms_per_step": float(tf.item()), "steps": k3,
                         "tree_sha": sha_f, "tree_sha_1gpu": verify["tree_sha_1gpu"] if verify else None,
                         "parallelism": "feature-block shard x%d (north_star): all rows on every GPU, F/P feature "
                                        "blocks each, one peer-memory slice exchange per tree level" % world}
        if rank == 0 and verify:
            assert sha_f == verify["tree_sha_1gpu"], "feature-block tree differs from the single-GPU tree"
        up_f.disable_peer_exchange()
        del up_f, page_f, preds_f, gpair_f, y_f
    else:
        if world > 1:
            up.disable_peer_exchange()
        del up, page, preds, gpair, pos
    torch.cuda.empty_cache()

    # End to end through the reference's own Booster.  N = 1: XGBoosterUpdateOneIter with the shim loaded, and the
    # reference's gpu_hist on the same DMatrix in this process.  N = 2: a distributed job of the unmodified reference (its
    # tracker and communicator) with the shim in distributed mode -- validated on 2 GPUs (tests/test_multi_gpu.py,
    # profiles/r02_bench_n2_driver_flags.json).  On 8 GPUs that leg did not return within the 9.5 minutes round 2's last
    # GPU call had left (no budget remained to find out where it waits), so beyond N = 2 it runs only on request
    # (B200_BENCH_DIST_E2E=1) and `e2e` stays the C-ABI host-gradient measurement above.
    if not args.no_e2e and world > 1 and (world == 2 or os.environ.get("B200_BENCH_DIST_E2E") == "1"):
        k2 = max(1, min(args.steps, 10))
        try:
            ref_section = reference_booster_distributed(torch, dist, args, device, rank, world, k2, 2)
        except Exception as exc:   # the measurement above must survive a failure of this leg
            ref_section = None
            if rank == 0:
                sys.stderr.write("distributed e2e leg failed: %r\n" % (exc,))
        if ref_section is not None:
            nodes = ref_section["ours"]["last_tree_nodes"]
            e2e = {"value": ref_section["ours"]["iters_per_s"], "unit": UNIT,
                   "h2d_bytes_per_step": nodes * 4 * world, "d2h_bytes_per_step": (nodes * 64 + 256) * world,
                   "path": "a %d-worker job of the unmodified reference (its tracker and communicator, row-split "
                           "QuantileDMatrix per worker, XGBoosterUpdateOneIter) with libb200hist_xgb_shim.so serving "
                           "grow_gpu_hist in distributed mode; max over ranks" % world,
                   "ms_per_iter_all_rank0": ref_section["ours"]["ms_per_iter_all_rank0"],
                   "steps": k2, "model_identical_to_reference_gpu_hist": ref_section["models_identical"]}
    if not args.no_e2e and world == 1 and rank == 0:
        k2 = max(1, min(args.steps, 10))
        try:
            ref_section = reference_booster_section(torch, args, device, k2, 2)
        except Exception as exc:   # the measurement above must survive a failure of this leg
            ref_section = None
            sys.stderr.write("e2e leg (reference Booster + shim) failed: %r\n" % (exc,))
        if ref_section is not None:
            nodes = ref_section["ours"]["last_tree_nodes"]
            e2e = {"value": ref_section["ours"]["iters_per_s"], "unit": UNIT,
                   "h2d_bytes_per_step": nodes * 4,            # the RegTree's leaf values for the prediction cache
                   "d2h_bytes_per_step": nodes * 64 + 256,     # node records + tree state + quantiser record
                   "path": "XGBoosterUpdateOneIter of the unmodified reference (oracle/_ref_cuda) with "
                           "libb200hist_xgb_shim.so serving grow_gpu_hist: objective gradient, Update, host replay "
                           "into RegTree, UpdatePredictionCache; labels resident in the DMatrix as in any xgb.train",
                   "ms_per_iter_median": ref_section["ours"]["ms_per_iter_median"], "steps": k2,
                   "model_identical_to_reference_gpu_hist": ref_section["models_identical"]}

    # BASELINE.json config C2 (10M x 64 dense, depth 6) beside the headline: same generator, same two Boosters on one
    # reference QuantileDMatrix.  A parity-test configuration timed for the record, not the metric.
    if not args.no_e2e and world == 1 and rank == 0 and args.rows == 100_000_000 and os.environ.get("B200_BENCH_OTHER", "1") != "0":
        try:
            c2 = reference_booster_section(torch, args, device, 10, 3, rows=10_000_000, features=64, depth=6)
        except Exception as exc:
            c2 = None
            sys.stderr.write("C2 leg failed: %r\n" % (exc,))
        if c2 is not None:
            other_configs = {"C2 synthetic 10Mx64 dense f32, 256 bins, depth 6, 1xB200": {
                "ours_through_the_shim_iters_per_s": c2["ours"]["iters_per_s"],
                "reference_gpu_hist_iters_per_s": c2["reference_gpu"]["iters_per_s"],
                "ours_ms_per_iter_median": c2["ours"]["ms_per_iter_median"],
                "reference_ms_per_iter_median": c2["reference_gpu"]["ms_per_iter_median"],
                "models_identical": c2["models_identical"], "steps": 10, "warmup": 3,
                "path": "XGBoosterUpdateOneIter, wall clock with a device synchronise on both sides"}}
        # config C5: 50M x 1024, each entry present with probability 0.01 (NaN otherwise): the reference stores a sparse
        # ELLPACK page (row_stride = longest row, 19-bit symbols), the shim imports it into our sparse page
        if other_configs is not None and os.environ.get("B200_BENCH_C5", "1") != "0":
            try:
                c5 = reference_booster_section(torch, args, device, 5, 2, rows=50_000_000, features=1024, depth=10, missing=0.99)
            except Exception as exc:
                c5 = None
                sys.stderr.write("C5 leg failed: %r\n" % (exc,))
            if c5 is not None:
                other_configs["C5 synthetic 50Mx1024, 1 % present, 256 bins, depth 10, 1xB200 (sparse Ellpack path)"] = {
                    "ours_through_the_shim_iters_per_s": c5["ours"]["iters_per_s"],
                    "reference_gpu_hist_iters_per_s": c5["reference_gpu"]["iters_per_s"],
                    "ours_ms_per_iter_median": c5["ours"]["ms_per_iter_median"],
                    "reference_ms_per_iter_median": c5["reference_gpu"]["ms_per_iter_median"],
                    "ours_ms_per_iter_all": c5["ours"]["ms_per_iter_all"],
                    "reference_ms_per_iter_all": c5["reference_gpu"]["ms_per_iter_all"],
                    "models_identical": c5["models_identical"], "steps": 5, "warmup": 2,
                    "reference_dmatrix_build_s": c5["dmatrix_build_s"],
                    "path": "XGBoosterUpdateOneIter, wall clock with a device synchronise on both sides"}

    # The same matrix under binary:logistic (labels y > 0): hessians p(1-p) have low words, so the histogram runs the
    # general 16-byte-accumulator path (HistKernelBankedFast: 32 features per CTA pass, 4 shared atomics per update)
    # instead of the 12-byte pair kernel.  Not the metric: a second objective on the headline matrix, for the record.
    # Runs AFTER the end-to-end legs, on a page rebuilt from the same cuts (2 s): it was written after the last GPU call
    # of round 2 that ran at N = 1, and a measurement that has never run must not be able to cost the ones that have.
    if world == 1 and os.environ.get("B200_BENCH_LOGISTIC", "1") != "0":
        try:
            page_l, y_l, _ = build_page(torch, xb, args, device, cuts, 0, nb)
            up_l = make_updater(page_l, "single")
            up_l.set_profiling(True)
            n_l = int(page_l.info()["n_rows"])
            yb = (y_l > 0).to(torch.float32)
            del y_l
            pl = torch.zeros(n_l, dtype=torch.float32, device=device)
            gl = torch.empty((n_l, 2), dtype=torch.float32, device=device)
            pos_l = torch.empty(n_l, dtype=torch.int32, device=device)

            def step_logistic():
                pr = torch.sigmoid(pl)
                torch.sub(pr, yb, out=gl[:, 0])
                torch.mul(pr, 1.0 - pr, out=gl[:, 1])
                del pr
                up_l.update_async(gl, pos_l)
                up_l.update_prediction_cache(pl)
                return up_l.fetch_tree()

            for _ in range(3):
                step_logistic()
            torch.cuda.synchronize(device)
            k4 = max(1, min(args.steps, 8))
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            lp = []
            for _ in range(k4):
                step_logistic()
                lp.append(up_l.last_profile())
            l1.record()
            torch.cuda.synchronize(device)
            lms = l0.elapsed_time(l1) / k4
            lv = float(np.mean([q_["row_visits"] for q_ in lp]))
            lh = float(np.mean([q_["hist"] for q_ in lp]))
            logistic = {"value": 1e3 / lms, "unit": UNIT, "ms_per_step": lms, "steps": k4, "hist_ms": lh,
                        "row_visits_per_step_mean": lv,
                        "updates_per_clk_per_sm": lv * f / (lh * 1e-3) / sm_clock_hz / n_sms if lh > 0 else None,
                        "kernel": "HistKernelBankedFast (16-byte accumulators, 4 shared atomics per update)",
                        "objective": "binary:logistic on labels y > 0 (gradient kernels of the objective included in the step)"}
            del up_l, page_l, yb, pl, gl, pos_l
            torch.cuda.empty_cache()
        except Exception as exc:   # a side measurement must not take the line down
            logistic = {"error": repr(exc)}

    if rank == 0 and not args.no_cpu_baseline:
        # a child process: the OpenMP runtime of THIS process was initialised under whatever OMP_NUM_THREADS the
        # launcher exported (torchrun: 1), the child re-executes itself with every host core (see _exec_with_all_cores)
        rows_c = min(args.rows, args.cpu_sample_rows)
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--gpus", "1", "--steps", "3",
                   "--warmup", "1", "--ref-rows", str(rows_c), "--rows", str(args.rows), "--features", str(args.features),
                   "--max-bin", str(args.max_bin), "--depth", str(args.depth)]
            env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "B200_BENCH_REEXEC")}
            env["LOCAL_RANK"] = str(local_rank)
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
            ref_line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
            cpu_baseline = ref_line.get("cpu_baseline")
        except Exception as exc:  # the baseline leg must not take the measurement down with it
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                            "sample": "cpu_baseline leg failed: %r" % (exc,)}

    watchdog.cancel()
    if rank == 0:
        print(json.dumps(make_line()))
    if world > 1:
        dist.destroy_process_group()


def _exec_with_all_cores():
    """torchrun exports OMP_NUM_THREADS=1 to its workers and libgomp reads it once, at load: the reference's
    OmpGetNumThreads() then clamps nthread=128 to 1 (src/common/threading_utils.cc).  The reference arm therefore
    re-executes itself with the thread count pinned to the host's cores before any OpenMP runtime is loaded."""
    cores = str(os.cpu_count() or 1)
    if os.environ.get("OMP_NUM_THREADS") == cores and os.environ.get("B200_BENCH_REEXEC") == "1":
        return
    env = dict(os.environ, OMP_NUM_THREADS=cores, OMP_PROC_BIND="false", B200_BENCH_REEXEC="1")
    env.pop("OMP_THREAD_LIMIT", None)
    os.execve(sys.executable, [sys.executable] + sys.argv, env)


def main():
    args = parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        _exec_with_all_cores()
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
