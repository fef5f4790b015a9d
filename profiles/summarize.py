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
    tim, rd, wr = collections.OrderedDict(), collections.defaultdict(float), collections.defaultdict(float)
    scale = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[start + 1:]:
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", "")) * scale.get(d.get("Metric Unit", ""), 1.0)
        k = d["Kernel Name"]
        if d["Metric Name"] == "gpu__time_duration.sum":
            tim.setdefault(k, []).append(v)
        elif d["Metric Name"] == "dram__bytes_read.sum":
            rd[k] += v
        elif d["Metric Name"] == "dram__bytes_write.sum":
            wr[k] += v
    tot = sum(sum(v) for v in tim.values())
    print(f"{'kernel':88s} {'n':>4s} {'avg us':>9s} {'total us':>10s} {'share':>6s} {'rd MB/launch':>13s} {'wr MB/launch':>13s} {'GB/s':>7s}")
    for k, v in sorted(tim.items(), key=lambda kv: -sum(kv[1])):
        n = len(v)
        r_, w_ = rd.get(k, 0.0) / n / 1e6, wr.get(k, 0.0) / n / 1e6
        bw = (r_ + w_) * 1e6 / (sum(v) / n) if (r_ + w_) else 0.0
        print(f"{k[:88]:88s} {n:4d} {sum(v) / n / 1e3:9.1f} {sum(v) / 1e3:10.1f} {100 * sum(v) / tot:5.1f}% {r_:13.1f} {w_:13.1f} {bw:7.0f}")


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
