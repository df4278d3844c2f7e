#!/usr/bin/env python
"""Turn ncu outputs (launch-list CSV, .ncu-rep) into the small text summaries kept under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv
    python profiles/summarize.py rep gpurun_out/prof.ncu-rep
"""
import collections
import csv
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "sm__cycles_elapsed.max")


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                v, u = float(d["Metric Value"].replace(",", "")), d["Metric Unit"]
                v = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
                name = d["Kernel Name"].split("(")[0]
                agg[name][0] += 1
                agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# per-kernel device time (ncu gpu__time_duration.sum, serialised, cold cache) from {path}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:28s} launches={v[0]:4d} total_us={v[1]:12.1f} avg_us={v[1] / v[0]:10.1f} share={100 * v[1] / tot:5.1f}%")


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"# kernel {d.get('Kernel Name')} grid {d.get('Grid Size')} block {d.get('Block Size')}  ({path})")
        for h, u, v in zip(hdr, units, r):
            if h in KEYS:
                print(f"{h:75s} {v:>18s} {u}")


if __name__ == "__main__":
    {"launches": launches, "rep": rep}[sys.argv[1]](sys.argv[2])
