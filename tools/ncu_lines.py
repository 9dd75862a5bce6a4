#!/usr/bin/env python
"""Aggregate warp-stall samples of an .ncu-rep per CUDA source line.
usage: tools/ncu_lines.py report.ncu-rep [launch_index] [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
# split in launches: a launch starts at the first "File Path" after a change of function block
blocks, cur, fname = [], None, None
seen_files = set()
for r in rows:
    if r and r[0] == 'File Path':
        f = r[1]
        if cur is None or f in seen_files:
            cur = {}
            blocks.append(cur)
            seen_files = set()
        seen_files.add(f)
        fname = f.split('/')[-1]
        continue
    if r and r[0] in ('Function Name', 'Line No'):
        if r[0] == 'Line No':
            hdr = r
            si = hdr.index('# Samples')
        continue
    if cur is None or not r or r[0] == '':
        continue
    try:
        n = int(r[si])
    except (ValueError, IndexError):
        continue
    key = (fname, int(r[0]), r[1].strip()[:100])
    cur[key] = cur.get(key, 0) + n
b = blocks[launch]
tot = sum(b.values())
print(f'launch {launch} of {len(blocks)}: {tot} samples')
for (f, ln, src), n in sorted(b.items(), key=lambda kv: -kv[1])[:topn]:
    print(f'{100.0 * n / tot:5.1f}%  {f}:{ln:<4d} {src}')
