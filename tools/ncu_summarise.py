#!/usr/bin/env python
"""Summarise an Nsight Compute report into the JSON shape kept under profiles/.

  ncu -i gpurun_out/<name>.ncu-rep --page raw --csv > /tmp/raw.csv      # here (no GPU needed)
  python tools/ncu_summarise.py /tmp/raw.csv profiles/<round>_<what>_ncu_summary.json "ncu command line" "note"

Per kernel name: launches captured, total / average duration, DRAM bytes (read + write) and GB/s, and the raw
metrics of the longest launch that DESIGN.md argues with (DRAM %, SM %, L1TEX data-stage %, issue %, stall ratios,
shared-atomic wavefronts, registers).  A launch list (`ncu --metrics gpu__time_duration.sum --csv`) goes through
`shares()` instead: kernel, launches, total ms, share of the step.
"""
import collections
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "smsp__inst_executed_op_shared_atom.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
MS = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3}


def short(name):
    return name.split("(")[0].replace("void ", "").replace("b200::", "")


def summarise(raw_csv):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in body:
        name = short(r[col["Kernel Name"]])
        ms = float(r[col["gpu__time_duration.sum"]]) * MS.get(units[col["gpu__time_duration.sum"]], 1)
        b = sum(float(r[col[k]]) * BYTES.get(units[col[k]], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        a = agg.setdefault(name, dict(launches=0, ms=0.0, bytes=0.0, longest=0.0, metrics=None))
        a["launches"] += 1
        a["ms"] += ms
        a["bytes"] += b
        if ms >= a["longest"]:
            a["longest"] = ms
            a["metrics"] = {k: r[col[k]] + " " + units[col[k]] for k in KEEP if k in col}
    return [dict(kernel=k, launches_captured=a["launches"], total_ms=round(a["ms"], 4),
                 avg_ms=round(a["ms"] / a["launches"], 4), dram_GB=round(a["bytes"] / 1e9, 4),
                 dram_GBps=round(a["bytes"] / a["ms"] / 1e6, 1) if a["ms"] > 0 else None,
                 longest_launch_metrics=a["metrics"]) for k, a in agg.items()]


def shares(launch_csv):
    rows = list(csv.reader(open(launch_csv)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mi:
            continue
        a = agg.setdefault(short(r[ki])[:70], [0, 0.0])
        a[0] += 1
        a[1] += float(r[mi].replace(",", "")) * MS.get(r[ui], 1)
    total = sum(a[1] for a in agg.values())
    return [(k, n, ms, ms / total) for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])]


if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    if sys.argv[1] == "--shares":
        with open(sys.argv[3], "w") as f:
            f.write("kernel,launches,total_ms,share\n")
            for k, n, ms, sh in shares(sys.argv[2]):
                f.write("%s,%d,%.3f,%.4f\n" % (k.replace(",", ";"), n, ms, sh))
    else:
        out = dict(command=sys.argv[3] if len(sys.argv) > 3 else "", note=sys.argv[4] if len(sys.argv) > 4 else "",
                   kernels=summarise(sys.argv[1]))
        json.dump(out, open(sys.argv[2], "w"), indent=1)
        for k in out["kernels"]:
            print(k["kernel"], k["launches_captured"], k["avg_ms"], "ms", k["dram_GBps"], "GB/s")
