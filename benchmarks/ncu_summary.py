#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` (read on stdin) into a small JSON: the metrics the roofline and
the issue analysis use, for the first captured launch of every kernel.
    ncu -i gpurun_out/r02_x.ncu-rep --page raw --csv | python benchmarks/ncu_summary.py > profiles/r02_x_ncu_summary.json"""
import csv
import json
import re
import sys

KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"lts__t_sector_hit_rate\.pct|l1tex__t_sector_hit_rate\.pct|sm__inst_executed\.(sum|avg\.per_cycle_elapsed)|"
                  r"sm__inst_executed_pipe_(alu|fma|fp64|xu|lsu|uniform)\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__(registers_per_thread|grid_size|block_size|occupancy_limit_\w+|shared_mem_per_block_\w+)|"
                  r"smsp__average_warps?_issue_stalled_\w+_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_\w+\.ratio|"
                  r"smsp__thread_inst_executed_per_inst_executed\.ratio|l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__cycles_elapsed\.max|smsp__warps_issue_stalled_\w+_per_warp_active\.pct)$")


def main():
    rows = list(csv.reader(sys.stdin))
    header = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            header, units, data = r, rows[i + 1], rows[i + 2:]
            break
    if header is None:
        raise SystemExit("no ncu raw csv on stdin")
    name_col = header.index("Kernel Name")
    out = {}
    for r in data:
        if len(r) != len(header):
            continue
        kname = re.sub(r"\(.*", "", r[name_col])
        if kname in out:
            continue
        out[kname] = {h: [r[j], units[j]] for j, h in enumerate(header) if KEEP.match(h)}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
