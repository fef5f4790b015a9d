#!/usr/bin/env python
"""Summarise ncu artefacts brought back from the GPU box (gpurun_out/) into small text files for profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv            # per-kernel launch-time table
    python profiles/summarize.py full gpurun_out/prof.ncu-rep                # key metrics of an `ncu --set full` capture
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 5]
    start = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr = rows[start]
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        if d.get("Metric Unit", "ns") in ("us", "usecond"):
            v *= 1e3
        elif d.get("Metric Unit", "ns") in ("ms", "msecond"):
            v *= 1e6
        agg.setdefault(d["Kernel Name"], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':90s} {'n':>5s} {'avg us':>10s} {'total us':>11s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:90]:90s} {len(v):5d} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / 1e3:11.1f} {100 * sum(v) / tot:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d["Kernel Name"], " grid", d.get("Grid Size"), " block", d.get("Block Size"))
        for k in KEYS:
            if k in d:
                print(f"  {k:82s} {d[k]:>22s} {units[hdr.index(k)]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
