#!/usr/bin/env python
"""Selected metrics of an `ncu --set full` report, one column per profiled launch.

    ncu -i prof.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv > profiles/rNN_ncu_full_summary.csv
"""
import csv
import re
import sys

KEEP = re.compile(
    r"^(Kernel Name|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|dram__bytes_(read|write)\.sum\.per_second|"
    r"dram__throughput\.avg\.pct|lts__t_sector_hit_rate\.pct|lts__t_bytes\.sum$|lts__throughput\.avg\.pct|"
    r"l1tex__data_pipe_lsu_wavefronts(_mem_shared)?\.(sum|avg)|l1tex__throughput\.avg\.pct|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|smsp__inst_executed\.sum$|smsp__issue_active\.avg\.pct|"
    r"sm__warps_active\.avg\.pct|sm__throughput\.avg\.pct|sm__inst_executed_pipe_(fma|fmaheavy|alu|lsu|uniform)\.sum$|"
    r"sm__pipe_fma_cycles_active\.avg\.pct|launch__(registers_per_thread|block_size|grid_size|shared_mem_per_block_dynamic|"
    r"occupancy_limit_\w+|waves_per_multiprocessor)$|sm__cycles_elapsed\.max$|"
    r"smsp__average_warps?_(latency_)?issue_stalled_\w+_per_issue_active)")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
for c, name in enumerate(hdr):
    if KEEP.match(name):
        w.writerow([name, units[c]] + [r[c] for r in data])
