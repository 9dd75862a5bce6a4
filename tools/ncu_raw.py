#!/usr/bin/env python
"""Print selected metrics of every launch in an .ncu-rep (raw page).
usage: tools/ncu_raw.py report.ncu-rep [extra_metric_substring ...]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__shared_mem_per_block_dynamic', 'l1tex__throughput.avg.pct_of_peak_sustained',
        'lts__throughput.avg.pct', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'launch__waves_per_multiprocessor', 'sm__pipe_fmaheavy', 'sm__inst_executed_pipe_lsu']
txt = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
want = WANT + sys.argv[2:]
cols = [i for i, h in enumerate(hdr) if any(w in h for w in want)]
ki, gi, bi = hdr.index('Kernel Name'), hdr.index('Grid Size'), hdr.index('Block Size')
for r in rows[2:]:
    print('==', r[ki][:60], 'grid', r[gi], 'block', r[bi])
    for i in cols:
        print(f'   {hdr[i].split(".", 2)[-1] if hdr[i].count(".") > 2 and hdr[i][0].isupper() else hdr[i]:75s} {r[i]:>16s} {units[i]}')
