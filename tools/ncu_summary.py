#!/usr/bin/env python
"""Summarise ncu output for profiles/.

  tools/ncu_summary.py launches <launches.csv>         -> per-kernel totals and shares
  tools/ncu_summary.py full <report.ncu-rep>           -> key metrics of every captured launch
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]


def launches(path):
    agg = collections.OrderedDict()
    n = 0
    for r in csv.reader(open(path)):
        if len(r) > 10 and r[0].isdigit():
            name = r[4].split("(")[0]
            v = float(r[-1].replace(",", ""))
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += v
            n += 1
    tot = sum(v[1] for v in agg.values())
    print("# %d launches, total %.3f ms (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)" % (n, tot / 1e6))
    print("%-80s %6s %12s %7s" % ("kernel", "n", "total_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-80s %6d %12.1f %6.1f%%" % (k[:80], v[0], v[1] / 1e3, 100 * v[1] / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("== %s  grid %s block %s" % (r[idx["Kernel Name"]][:100], r[idx["Grid Size"]], r[idx["Block Size"]]))
        for k in KEYS:
            if k in idx:
                print("   %-72s %s %s" % (k, r[idx[k]], units[idx[k]]))
        st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(r[i].replace(",", "") or 0)) for h, i in idx.items()
              if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
        tot = sum(v for _, v in st) or 1.0
        print("   stall samples: " + ", ".join("%s %.1f%%" % (h, 100 * v / tot) for h, v in sorted(st, key=lambda x: -x[1])[:8]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
