#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files that
are committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
  python profiles/summarize.py full     gpurun_out/prof.ncu-rep  > profiles/rNN_full.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

FULL_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    name_i, val_i = hdr.index("Kernel Name"), hdr.index("Metric Value")
    unit_i = hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rows[1:]:
        if len(r) <= val_i:
            continue
        v = float(r[val_i].replace(",", ""))
        u = r[unit_i]
        ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        k = r[name_i].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += ms
        total += ms
    print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:8d} {ms:10.3f} {ms / n:9.4f} {100 * ms / total:6.1f}%")
    print(f"{'total':70s} {'':8s} {total:10.3f}")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print(r[hdr.index("Kernel Name")])
        for k in FULL_KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:62s} {r[i]:>18s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h))
                except ValueError:
                    pass
        print("  top stall reasons (warps stalled per issue-active cycle):")
        for v, h in sorted(stalls, reverse=True)[:6]:
            short = h.replace("smsp__average_warps_issue_stalled_", "").replace(
                "_per_issue_active.ratio", "")
            print(f"    {short:30s} {v:8.2f}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
