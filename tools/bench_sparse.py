#!/usr/bin/env python
"""Config C5 of BASELINE.json / SURVEY.md 8(d): synthetic CSR matrix (each entry present w.p. `density`, present
values N(0,1)), sparse ELLPACK path (uint32 global-bin symbols, row_stride = longest row), depth 10.

Not the bench.py headline: a parity-test configuration timed for the record.
  python tools/bench_sparse.py --rows 50000000 --features 1024 --density 0.01 --depth 10 --steps 5 --warmup 2
Prints one JSON line: iters/s, per-phase device times, page geometry, DMatrix-build seconds.  Batches of 1M rows are
generated on the device; cuts come from our quantile sketch fed with the dense image of every batch (NaN = absent).
With --reference the same stream also goes through the UNMODIFIED reference (oracle/_ref_cuda) for a side-by-side.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gen_batch(torch, dev, b, rows, f, density):
    g = torch.Generator(device=dev).manual_seed(11 + b)
    present = torch.rand((rows, f), generator=g, device=dev) < density
    x = torch.randn((rows, f), generator=g, device=dev)
    x = torch.where(present, x, torch.full_like(x, float("nan")))
    return x, present, g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=50_000_000)
    ap.add_argument("--features", type=int, default=1024)
    ap.add_argument("--density", type=float, default=0.01)
    ap.add_argument("--max-bin", type=int, default=256)
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=1_000_000)
    ap.add_argument("--reference", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    import xgboost_b200 as xb
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n, f = args.rows, args.features
    nb = (n + args.batch - 1) // args.batch

    # pass 1: sketch + longest row
    t0 = time.perf_counter()
    sk = xb.QuantileSketch(f, args.max_bin, dev)
    stride = 1
    for b in range(nb):
        rows = min(args.batch, n - b * args.batch)
        x, present, _ = gen_batch(torch, dev, b, rows, f, args.density)
        sk.push(x)
        stride = max(stride, int(present.sum(dim=1).max().item()))
        del x, present
    cuts = sk.make_cuts()
    sk.close()
    t_sketch = time.perf_counter() - t0

    # pass 2: CSR of every batch -> sparse page
    t0 = time.perf_counter()
    page = xb.EllpackPage(cuts, n, 0, sparse_row_stride=stride)
    w = torch.randn(f, generator=torch.Generator(device=dev).manual_seed(7), device=dev)
    y = torch.empty(n, dtype=torch.float32, device=dev)
    nnz_total = 0
    for b in range(nb):
        rows = min(args.batch, n - b * args.batch)
        x, present, g = gen_batch(torch, dev, b, rows, f, args.density)
        counts = present.sum(dim=1)
        indptr = torch.zeros(rows + 1, dtype=torch.int64, device=dev)
        indptr[1:] = torch.cumsum(counts, 0)
        idx = present.nonzero(as_tuple=False)             # row-major order = CSR order
        indices = idx[:, 1].to(torch.int32).contiguous()
        values = x[present].contiguous()
        page.push_csr(indptr, indices, values, b * args.batch)
        y[b * args.batch:b * args.batch + rows] = torch.nan_to_num(x, nan=0.0) @ w + \
            0.1 * torch.randn(rows, generator=g, device=dev)
        nnz_total += int(indices.numel())
        del x, present, idx, indices, values
    torch.cuda.synchronize()
    t_page = time.perf_counter() - t0

    up = xb.GPUHistMaker(page)
    up.configure(dict(max_depth=args.depth, eta=0.3, reg_lambda=1.0, min_child_weight=1.0))
    preds = torch.zeros(n, dtype=torch.float32, device=dev)
    gpair = torch.empty((n, 2), dtype=torch.float32, device=dev)
    gpair[:, 1] = 1.0
    pos = torch.empty(n, dtype=torch.int32, device=dev)

    def step():
        torch.sub(preds, y, out=gpair[:, 0])
        tree = up.update(gpair, pos)
        up.update_prediction_cache(preds)
        return tree

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tree = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    up.set_profiling(True)
    step()
    prof = up.last_profile()
    up.set_profiling(False)
    line = {"metric": "boosting iters/sec, %dx%d CSR %.3g dense, %d bins, depth %d (sparse ELLPACK path)" %
                      (n, f, args.density, args.max_bin, args.depth),
            "value": 1e3 / ms, "unit": "iters/s", "ms_per_step": ms, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "row_stride": stride, "nnz": nnz_total, "total_bins": int(cuts.total_bins),
            "tree_nodes": tree.num_nodes, "phase_ms": {k: v for k, v in prof.items() if k != "row_visits"},
            "row_visits": prof["row_visits"], "dmatrix_build_s": {"sketch": round(t_sketch, 2), "page": round(t_page, 2)},
            "mse": float(((preds - y) ** 2).mean().item())}
    if args.reference:
        from oracle import ref_capi as rc
        batches = rc.LazyBatches(nb, lambda b: gen_batch(torch, dev, b, min(args.batch, n - b * args.batch), f,
                                                         args.density)[0])
        t0 = time.perf_counter()
        dm = rc.DMatrix.quantile_from_batches(batches, max_bin=args.max_bin, kind="cuda")
        dm.set_label(y.cpu().numpy())
        t_build = time.perf_counter() - t0
        bst = rc.Booster({"tree_method": "hist", "device": "cuda:0", "max_depth": args.depth, "max_bin": args.max_bin,
                          "objective": "reg:squarederror", "eta": 0.3, "lambda": 1, "min_child_weight": 1}, dm)
        for i in range(args.warmup):
            bst.update(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.steps):
            bst.update(args.warmup + i)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        line["reference_gpu_hist"] = {"iters_per_s": 1.0 / dt, "ms_per_iter": dt * 1e3, "dmatrix_build_s": t_build}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
