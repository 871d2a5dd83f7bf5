#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small text files under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r01.csv > profiles/r01_launches.txt
  python profiles/summarize.py kernel   gpurun_out/prof_r01.ncu-rep  > profiles/r01_pa_apply_full.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hd = rows[h]
    ki, vi = hd.index("Kernel Name"), hd.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[h + 1:]:
        if len(r) > vi:
            name = r[ki].split("(")[0][:80]
            tot[name] += float(r[vi].replace(",", ""))
            cnt[name] += 1
    s = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {s / 1e6:.3f} ms total (gpu__time_duration.sum, "
          "--clock-control none; cold-cache, serialised: compare shares, not absolutes)")
    print(f"{'total_us':>12} {'count':>6} {'avg_us':>10} {'share':>7}  kernel")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v / 1e3:12.1f} {cnt[k]:6d} {v / 1e3 / cnt[k]:10.1f} {100 * v / s:6.1f}%  {k}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], check=True,
                         stdout=subprocess.PIPE).stdout.decode()
    rows = list(csv.reader(out.splitlines()))
    hd, units, data = rows[0], rows[1], rows[2:]
    print(f"# {path}: ncu --set full --clock-control none, {len(data)} captured launch(es)")
    ki = hd.index("Kernel Name")
    for r in data:
        print("kernel:", r[ki][:120])
    for k in KEYS:
        if k in hd:
            i = hd.index(k)
            print(f"{k:85s} {units[i]:14s} " + "  ".join(r[i] for r in data))
    print("# warp stall reasons, first launch")
    for i, h in enumerate(hd):
        if "issue_stalled" in h and "not_issued" not in h and ("ratio" in h or "pct" in h):
            print(f"{h:85s} {units[i]:14s} {data[0][i]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
