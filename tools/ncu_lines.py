#!/usr/bin/env python
"""Per CUDA source line: stall samples and executed warp-instructions per kernel.
usage: tools/ncu_lines.py report.ncu-rep [kernel_substring] [top_n] [sort: samples|inst]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ''
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
key = sys.argv[4] if len(sys.argv) > 4 else 'samples'
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
kern = {}      # function name -> {(file, line, src): [samples, inst]}
fname = func = hdr = None
ln, src = None, None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        func = r[1]
        continue
    if r[0] == 'Line No':
        hdr = r
        si, ei = hdr.index('# Samples'), hdr.index('Instructions Executed')
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] != '':
        ln, src = r[0], r[1].strip()[:95]
    try:
        ns, ne = int(r[si] or 0), int(r[ei] or 0)
    except ValueError:
        continue
    d = kern.setdefault(func, {})
    a = d.setdefault((fname, ln, src), [0, 0])
    a[0] += ns
    a[1] += ne
for func, b in kern.items():
    if want not in func:
        continue
    # every launch of the same function is listed again: totals are summed over them
    ts, te = sum(v[0] for v in b.values()), sum(v[1] for v in b.values())
    print(f'== {func}: {ts} samples, {te} warp-instructions (summed over captured launches)')
    idx = 0 if key == 'samples' else 1
    for (f, l, s), v in sorted(b.items(), key=lambda kv: -kv[1][idx])[:topn]:
        print(f'{100.0 * v[0] / max(ts, 1):5.1f}% smp {100.0 * v[1] / max(te, 1):5.1f}% inst  {f}:{l:<4s} {s}')
