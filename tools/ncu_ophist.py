#!/usr/bin/env python
"""Executed-opcode histogram of a SASS index range of kernel 0 in an .ncu-rep.
usage: tools/ncu_ophist.py report.ncu-rep first last [frames]"""
import collections
import csv
import subprocess
import sys

rep, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
frames = float(sys.argv[4]) if len(sys.argv) > 4 else 5920.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
k = [r for r in csv.reader(out.splitlines()) if r and r[0].startswith('0x')]
c, tot = collections.Counter(), 0
for i in range(a, min(b + 1, len(k))):
    t = k[i][1].strip()
    op = t.split()[1] if t.startswith('@') else t.split()[0]
    c[op.split('.')[0]] += int(k[i][5])
    tot += int(k[i][5])
print('sass %d-%d: %d warp-instructions, %.0f per frame' % (a, b, tot, tot / frames))
for o, v in c.most_common(30):
    print('   %-10s %5.1f%%  %.0f/frame' % (o, 100.0 * v / tot, v / frames))
