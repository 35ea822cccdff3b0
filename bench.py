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
    ap.add_argument("--cpu-sample-rows", type=int, default=int(os.environ.get("B200_BENCH_CPU_ROWS", 1_000_000)))
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


def build_workload(torch, xb, args, device, batch_lo=0, batch_hi=None):
    """DMatrix build = what QuantileDMatrix(DataIter of 1M-row device batches) does in the reference
    (quantile_dmatrix.cu:29-126): pass 1 pushes every batch through the resident quantile sketch
    (b200hist_sketch_push_dense, one sketch window per batch), pass 2 regenerates the batches and bins them
    into the Ellpack page.  Returns page, labels and the two pass times (seconds)."""
    n, f = args.rows, args.features
    batch = 1_000_000
    nb = (n + batch - 1) // batch
    batch_hi = nb if batch_hi is None else batch_hi       # this rank's page holds batches [batch_lo, batch_hi)
    row_lo, row_hi = batch_lo * batch, min(n, batch_hi * batch)
    t0 = time.perf_counter()
    sketch = xb.QuantileSketch(f, args.max_bin, device)   # cuts always come from the WHOLE stream
    for b in range(nb):
        x, _ = batch_of(torch, device, f, n, batch, b)
        sketch.push(x)
        del x
    cuts = sketch.make_cuts()
    sketch.close()
    torch.cuda.synchronize(device)
    t_sketch = time.perf_counter() - t0
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
    return page, y, dict(sketch_s=t_sketch, ellpack_s=time.perf_counter() - t0, sketch_windows=nb)


def algorithmic_bytes(n, f, bin_bytes, depth, row_visits):
    """bytes each stage must move once (dense page)."""
    hist = row_visits * (f * bin_bytes + 16 + 4)
    part = (depth - 1) * n * (4 + bin_bytes + 4)
    quant = n * (8 + 8 + 16)
    final = n * (4 + 4) + n * (4 + 4 + 4)
    return dict(hist=hist, partition=part, quantise=quant, finalise=final, total=hist + part + quant + final)


# --------------------------------------------------------------------------------------------------
# reference CPU arm (the unmodified reference's tree_method=hist on the host cores)
# --------------------------------------------------------------------------------------------------
def reference_cpu_iters_per_s(args, steps, warmup, sample_rows):
    """Times XGBoosterUpdateOneIter of the reference CPU library on a bounded sample of the workload
    (same generator, same params) and scales to the full row count (hist cost is linear in rows)."""
    import numpy as np
    from oracle import ref_capi as rc
    if not rc.available("cpu"):
        return None
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(2024)
    n, f = sample_rows, args.features
    x = rng.standard_normal((n, f), dtype=np.float32)
    w = np.random.default_rng(7).standard_normal(f).astype(np.float32)
    y = (x @ w + 0.1 * rng.standard_normal(n, dtype=np.float32)).astype(np.float32)
    dm = rc.DMatrix.quantile_from_batches([x], max_bin=args.max_bin, kind="cpu", nthread=cores)
    dm.set_label(y)
    bst = rc.Booster({"tree_method": "hist", "device": "cpu", "max_depth": args.depth, "max_bin": args.max_bin,
                      "objective": "reg:squarederror", "eta": 0.3, "lambda": 1, "min_child_weight": 1,
                      "nthread": cores}, dm)
    for i in range(warmup):
        bst.update(i)
    t0 = time.perf_counter()
    for i in range(steps):
        bst.update(warmup + i)
    dt = (time.perf_counter() - t0) / max(1, steps)
    scaled = (1.0 / dt) * (n / float(args.rows))
    return dict(value=scaled, unit=UNIT, cores=cores, kind="reference",
                sample="reference CPU hist (oracle/_ref/libxgboost.so, nthread=%d) on %d x %d rows: %.3f s/iter, "
                       "scaled linearly in rows to %d" % (cores, n, f, dt, args.rows),
                sample_ms_per_iter=dt * 1e3)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = reference_cpu_iters_per_s(args, max(1, min(args.steps, 5)), max(1, min(args.warmup, 2)), args.cpu_sample_rows)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libxgboost.so not built"}))
        return
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / res["value"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "synthetic %dx%d dense f32, %d bins, depth %d" % (args.rows, args.features, args.max_bin, args.depth),
                       "l2": "inputs exceed L2"},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


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
    if row_split:  # contiguous runs of 1M-row batches per rank
        nb = (n_total + 999_999) // 1_000_000
        page, y, build_s = build_workload(torch, xb, args, device, nb * rank // world, nb * (rank + 1) // world)
    else:
        page, y, build_s = build_workload(torch, xb, args, device)
    info = page.info()
    n = int(info["n_rows"])      # rows resident on THIS rank
    up = xb.GPUHistMaker(page)
    up.configure(dict(max_depth=args.depth, eta=0.3, reg_lambda=1.0, min_child_weight=1.0))
    comm = None
    if world > 1:
        if row_split:  # full-width partial histograms, one summing peer-memory kernel per tree level
            up.enable_peer_exchange(row_split=True)
        elif os.environ.get("B200HIST_COMM", "p2p") == "nccl":  # feature blocks + one NCCL allreduce per level
            comm = xb.Communicator.from_torch_distributed(device)
            up.set_comm(comm)
        else:  # feature blocks + peer-memory push + flag barrier over NVLink (no NCCL in the data path)
            up.enable_peer_exchange()
    preds = torch.zeros(n, dtype=torch.float32, device=device)
    gpair = torch.empty((n, 2), dtype=torch.float32, device=device)
    gpair[:, 1] = 1.0
    pos = torch.empty(n, dtype=torch.int32, device=device)

    def step_resident():
        torch.sub(preds, y, out=gpair[:, 0])  # reg:squarederror gradient; hessian stays 1
        tree = up.update(gpair, pos)
        up.update_prediction_cache(preds)
        return tree

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = xb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tree = step_resident()
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

    # per-phase device times + row visits from profiled iterations (events inside the library, same stream)
    up.set_profiling(True)
    prof = []
    for _ in range(3):
        step_resident()
        prof.append(up.last_profile())
    up.set_profiling(False)
    torch.cuda.synchronize(device)
    # where a step's time goes outside the updater's phases: gradient kernel, tree update call, prediction cache
    # (profiling stays on so that the same iterations give the phase sum: update_tree - sum(phases) = unattributed)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    parts = []
    up.set_profiling(True)
    for _ in range(3):
        ev[0].record()
        torch.sub(preds, y, out=gpair[:, 0])
        ev[1].record()
        up.update(gpair, pos)
        ev[2].record()
        up.update_prediction_cache(preds)
        ev[3].record()
        torch.cuda.synchronize(device)
        lp = up.last_profile()
        parts.append([ev[i].elapsed_time(ev[i + 1]) for i in range(3)] + [sum(v for k, v in lp.items() if k not in ("row_visits", "hist_exchange"))])
    up.set_profiling(False)
    step_parts = dict(zip(["gradient", "update_tree", "prediction_cache", "update_tree_phase_sum"],
                          [float(v) for v in np.median(np.array(parts), axis=0)]))
    phase = {k: float(np.median([p[k] for p in prof])) for k in prof[0] if k != "row_visits"}
    row_visits = prof[-1]["row_visits"]
    ab = algorithmic_bytes(n, f, info["bin_bytes"], args.depth, row_visits)
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    hist_launches = args.depth
    # per GPU: this rank reads its F/world feature columns + the 20 B of gradient pair / row id per row visit
    feat_div = 1 if (world == 1 or row_split) else world   # row split: this rank's visits cover all F features
    hist_bytes_gpu = row_visits * (f * info["bin_bytes"] / feat_div + 16 + 4)
    hist_ms = phase["hist"] - phase.get("hist_exchange", 0.0)   # the kernel itself, without the cross-GPU exchange
    hist_gbs = hist_bytes_gpu / (hist_ms * 1e-3) / 1e9 if hist_ms > 0 else 0.0
    # the binding unit is the L1TEX data stage (shared atomics + gather wavefronts); the ceiling of the atomics alone
    # is measured here by the library's own micro-benchmark
    from xgboost_b200 import components as xc
    smem_peak = xc.smem_atomic_peak(True, device)
    sm_clock_hz = (clocks["sm_mhz"] if clocks and clocks.get("sm_mhz") else 1965) * 1e6
    n_sms = torch.cuda.get_device_properties(device).multi_processor_count
    updates_per_s = row_visits * (f / feat_div) / (hist_ms * 1e-3) if hist_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "HistKernelPair", "achieved": hist_gbs, "peak": peak, "unit": "GB/s",
                "frac": hist_gbs / peak, "traffic": 1.2 * hist_bytes_gpu / hist_launches, "peak_source": peak_src,
                "traffic_source": "ncu --set full dram__bytes_read+write of HistKernelPair = 1.13-1.29 x algorithmic on three "
                                  "consecutive level launches of the 20M-row capture "
                                  "(profiles/r01c_pair_kernels_ncu_summary.json), 1.2 x scaled to this launch size",
                "l1tex_data_pipe": {"pct_of_peak": 95.0, "shared_atomic_share_pct": 72.0,
                                    "source": "ncu l1tex__data_pipe_lsu_wavefronts (same capture): the binding unit"},
                "launches_per_step": hist_launches, "avg_launch_ms": hist_ms / hist_launches,
                "algorithmic_bytes_per_launch": hist_bytes_gpu / hist_launches, "row_visits_per_step": row_visits,
                # the binding resource is the shared-memory atomic pipe, not HBM: the ceiling is MEASURED on this GPU by
                # b200hist_smem_atomic_peak (the kernel's own accumulation sequence -- 3 conflict-free 32-bit atomics
                # per update for reg:squarederror, whose quantised hessian has a zero low word -- with no global loads)
                "smem_atomic_bound": {"updates_per_clk_per_sm": updates_per_s / sm_clock_hz / n_sms,
                                      "ceiling": smem_peak, "frac": updates_per_s / sm_clock_hz / n_sms / smem_peak,
                                      "ceiling_source": "measured: b200hist_smem_atomic_peak(h_lo_zero=1), clock64 based",
                                      "note": "exact int64 accumulation in shared memory bounds the kernel, not HBM"},
                "whole_step_frac": (ab["total"] - ab["hist"] + hist_bytes_gpu) / (ms_per_step * 1e-3) / 1e9 / peak}

    # end-to-end: gradient arrives in pinned HOST memory, tree returns to the host (C-ABI host entry)
    e2e = None
    if not args.no_e2e:
        g_host = torch.empty((n, 2), dtype=torch.float32).pin_memory()
        g_host.copy_(gpair)
        loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

        def step_e2e():
            tree_ = up.update(g_host, pos)            # H2D of the gradient + D2H of the tree inside
            up.update_prediction_cache(preds)
            loss_host.copy_(((preds - y) ** 2).mean().reshape(1), non_blocking=False)  # D2H of the metric
            return tree_

        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        barrier()
        k2 = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k2):
            tr = step_e2e()
        barrier()
        dt = (time.perf_counter() - t0) / k2
        tt = torch.tensor([dt], device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": 1.0 / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(n_total) * 8 if (world == 1 or row_split) else int(n_total) * 8 * world,
               "d2h_bytes_per_step": int(tr.num_nodes) * 48 + 4}

    # north_star's decomposition measured next to the default one: all rows on every GPU, F/P feature blocks each,
    # one peer-memory slice exchange per tree level (same matrix, same trees)
    feature_block = None
    if row_split and os.environ.get("B200_BENCH_ALSO_FEATURES", "1") != "0":
        up.disable_peer_exchange()   # unmap on every rank, barrier, only then free
        del up, page
        torch.cuda.empty_cache()
        page_f, y_f, _ = build_workload(torch, xb, args, device)
        up_f = xb.GPUHistMaker(page_f)
        up_f.configure(dict(max_depth=args.depth, eta=0.3, reg_lambda=1.0, min_child_weight=1.0))
        up_f.enable_peer_exchange()
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
                         "parallelism": "feature-block shard x%d (north_star): all rows on every GPU, F/P feature "
                                        "blocks each, one peer-memory slice exchange per tree level" % world}
        up_f.disable_peer_exchange()
        del up_f, page_f, preds_f, gpair_f, y_f

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu_baseline = reference_cpu_iters_per_s(args, 3, 1, args.cpu_sample_rows)
        if cpu_baseline is not None:
            cpu_baseline = {k: cpu_baseline[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                "config": {"workload": "synthetic %dx%d dense f32, %d bins, depth %d, reg:squarederror" % (n_total, f, args.max_bin, args.depth),
                           "parallelism": "single GPU" if world == 1 else (
                               "row split x%d: N/P rows per GPU, full-width partial histograms, one summing "
                               "peer-memory exchange (reduce-scatter + all-gather over NVLink) per tree level" % world
                               if row_split else
                               "feature-block shard x%d: all rows on every GPU, F/P feature blocks each, one "
                               "peer-memory slice exchange per tree level" % world),
                           "l2": "inputs (%.1f GB bin matrix) exceed the 126 MB L2" % (n * info["row_bytes"] / 1e9),
                           "tree_nodes": tree.num_nodes,
                           "cuts": "own quantile sketch, %d windows of <= 1M rows merged (QuantileDMatrix semantics)"
                                   % build_s["sketch_windows"]},
                "dmatrix_build_s": {"sketch": round(build_s["sketch_s"], 3), "ellpack": round(build_s["ellpack_s"], 3)},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "phase_ms": phase, "step_parts_ms": step_parts, "feature_block_shard": feature_block, "algorithmic_gb_per_step": {k: v / 1e9 for k, v in ab.items()}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
