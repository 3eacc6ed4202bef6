#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key counters of the kernel and the
per-source-line share of executed instructions / stall samples.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [ntop]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for val in rows[2:]:
        print("kernel:", val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
        for i, h in enumerate(hdr):
            if h in KEYS or ("issue_stalled" in h and "per_warp_active" in h and float(val[i] or 0) > 3):
                print(f"  {h:90s} {val[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = None
    cur = None
    fname = None
    agg = {}
    ti = ts = 0
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 8:
            continue
        if r[0] != "":
            cur = (fname, int(r[0]), r[1].strip()[:100])
            agg.setdefault(cur, [0, 0])
        else:
            try:
                sa, ins = int(r[4]), int(r[7])
            except ValueError:
                continue
            agg[cur][0] += sa
            agg[cur][1] += ins
            ti += ins
            ts += sa
    print(f"total warp instructions {ti}, stall samples {ts}")
    byfile = {}
    for (f, l, s), (sa, i) in agg.items():
        byfile.setdefault(f, [0, 0])
        byfile[f][0] += sa
        byfile[f][1] += i
    for f, (sa, i) in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
        if i or sa:
            print(f"  {f:34s} inst {100 * i / max(ti, 1):5.1f}%  samples {100 * sa / max(ts, 1):5.1f}%")
    print("top lines by stall samples:")
    for (f, l, s), (sa, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:ntop]:
        print(f"  {f}:{l:<4d} inst {100 * i / max(ti, 1):5.1f}% samp {100 * sa / max(ts, 1):5.1f}%  {s}")


if __name__ == "__main__":
    main()
