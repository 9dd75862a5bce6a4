#!/usr/bin/env python
"""Per-launch summary of an .ncu-rep's SASS page: stall-reason totals,
executed-instruction mix by opcode, and the instructions with most stall samples.
usage: tools/ncu_sass.py report.ncu-rep [launch_index] [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
for li, s in enumerate(starts):
    if which is not None and li != which:
        continue
    e = starts[li + 1] if li + 1 < len(starts) else len(rows)
    h = rows[s]
    body = [r for r in rows[s + 1:e] if len(r) == len(h)]
    ci = {n: h.index(n) for n in h}
    stall_cols = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
    tot = collections.Counter()
    mix = collections.Counter()
    samples = []
    for r in body:
        ns = int(r[ci['# Samples']] or 0)
        ex = int(r[ci['Instructions Executed']] or 0)
        op = r[ci['Source']].split()
        op = [o for o in op if not o.startswith('@')]
        opc = op[0].split('.')[0] if op else '?'
        mix[opc] += ex
        for c in stall_cols:
            tot[c] += int(r[ci[c]] or 0)
        top = max(stall_cols, key=lambda c: int(r[ci[c]] or 0))
        samples.append((ns, r[ci['Source']].strip()[:70], top, ex))
    nsamp = sum(x[0] for x in samples)
    nex = sum(mix.values())
    print(f'== launch {li}: {len(body)} SASS lines, {nsamp} samples, {nex} warp-instructions')
    print('   stalls: ' + ', '.join(f'{k[6:]} {100*v/max(1,sum(tot.values())):.1f}%'
                                    for k, v in tot.most_common(9)))
    print('   mix:    ' + ', '.join(f'{k} {100*v/max(1,nex):.1f}%' for k, v in mix.most_common(18)))
    for ns, src, top, ex in sorted(samples, reverse=True)[:topn]:
        print(f'   {100*ns/max(1,nsamp):5.1f}%  {src:70s} {top[6:]:10s} x{ex}')
