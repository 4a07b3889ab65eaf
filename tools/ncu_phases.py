#!/usr/bin/env python
"""Split a kernel's SASS at BAR.SYNC instructions and report executed warp-instructions per region
(regions appear in program order = kernel phases).  usage: tools/ncu_phases.py report.ncu-rep [kernel#]"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', str(which),
                      '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kernels, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = []
        kernels.append(cur)
    elif cur is not None and r and r[0].startswith('0x'):
        cur.append(r)
k = kernels[0]
name = next((r[1] for r in rows if r and r[0] == 'Kernel Name'), '?').split('(b200phy')[0]
total = sum(int(r[5]) for r in k)
print('kernel %d %s: %d SASS instructions, %d executed warp-instructions' % (which, name, len(k), total))
seg, ops, n0 = 0, Counter(), 0
def flush(i, tag):
    global seg, ops, n0
    if seg:
        top = ', '.join('%s %.0f%%' % (o, 100.0 * c / seg) for o, c in ops.most_common(6))
        print('  sass %5d-%5d %6.2f%%  %s  | %s' % (n0, i, 100.0 * seg / total, tag, top))
    seg, ops, n0 = 0, Counter(), i + 1
for i, r in enumerate(k):
    txt = r[1].strip()
    op = txt.split()[0] if not txt.startswith('@') else txt.split()[1]
    op = op.split('.')[0]
    e = int(r[5])
    seg += e
    ops[op] += e
    if op == 'BAR':
        flush(i, 'BAR')
flush(len(k), 'END')
