#!/usr/bin/env python
"""Dump kernel 0 of an .ncu-rep as `index  executed-per-frame  SASS` lines (whole kernel or a range).
usage: tools/ncu_sass_dump.py report.ncu-rep [first last] [frames]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
a = int(sys.argv[2]) if len(sys.argv) > 3 else 0
b = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
frames = float(sys.argv[4]) if len(sys.argv) > 4 else 5920.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
k = [r for r in csv.reader(out.splitlines()) if r and r[0].startswith('0x')]
for i in range(a, min(b + 1, len(k))):
    print('%5d %8.2f  %s' % (i, int(k[i][5]) / frames, k[i][1].strip()))
