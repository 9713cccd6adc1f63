#!/usr/bin/env python
"""Summarise ncu output for profiles/ (run here, no GPU needed).

  tools/ncu_summary.py raw   <file.ncu-rep> <out.csv>     selected metrics per profiled launch
  tools/ncu_summary.py list  <launches.csv> <out.txt>     per-kernel count / mean / share of a launch list
"""
import collections
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct"]


def raw(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])


def launch_list(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > vi:
            agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# per-kernel summary of {path} (gpu__time_duration.sum, ns; cold-cache serialised launches: compare SHARES)\n")
        for k, v in agg.items():
            f.write(f"{k:60s} launches={len(v):5d} mean_us={sum(v) / len(v) / 1e3:10.2f} share={sum(v) / tot * 100:6.2f}%\n")


if __name__ == "__main__":
    {"raw": raw, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
