#!/usr/bin/env python
"""Turns an .ncu-rep (one kernel, `ncu --set full`) into the short text summary kept under profiles/:
    python tools/ncu_summary.py gpurun_out/x.ncu-rep "how it was captured" > profiles/x_ncu_summary.txt
and an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals:
    python tools/ncu_summary.py --launches gpurun_out/launches.csv > profiles/launches_summary.txt"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

KEEP = re.compile(r"^(gpu__time_duration\.sum|launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_.*)|"
                  r"dram__bytes_(read|write)\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?|lts__throughput.*elapsed|l1tex__throughput.*elapsed|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|sm__throughput.*elapsed|sm__warps_active.*|smsp__issue_active.*active|smsp__inst_executed\.sum|"
                  r"sm__pipe_(tensor|fma|fmaheavy|alu|fp64)\w*cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
                  r"sm__inst_executed_pipe_(alu|fma|fmaheavy|lsu|tma|tensor\w*|uniform)\.(avg|sum)\.pct_of_peak_sustained_active|sm__cycles_elapsed\.max|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio)$")


def summary(rep, how):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print(how)
    print("kernel:", vals[hdr.index("Kernel Name")], "\n")
    for i, name in enumerate(hdr):
        if KEEP.match(name) and not re.search(r"\.(max|min)\.|\.sum\.(pct|per_cycle)|per_cycle_active", name):
            try:
                if "stalled" in name and float(vals[i]) < 0.5:
                    continue
            except ValueError:
                pass
            print(f"{name:90s} {units[i]:16s} {vals[i]}")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    t, n = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        t[r[ki][:96]] += v
        n[r[ki][:96]] += 1
    tot = sum(t.values())
    print(f"total GPU time of the process: {tot / 1e6:.3f} ms over {sum(n.values())} launches (ncu: serialised, cold cache)")
    for k, v in sorted(t.items(), key=lambda x: -x[1])[:14]:
        print(f"{v / 1e6:10.3f} ms {100 * v / tot:5.1f} %  n={n[k]:4d}  avg {v / n[k] / 1e6:8.3f} ms  {k}")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        summary(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
