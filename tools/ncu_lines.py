#!/usr/bin/env python
"""Summarise an .ncu-rep per CUDA source line: instructions executed and stall samples.
usage: tools/ncu_lines.py report.ncu-rep [top_n [launch#]]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = sys.argv[3] if len(sys.argv) > 3 else '0'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--launch-skip', which, '--launch-count', '1'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, res, seen_kernel = None, [], 0
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if r[0].isdigit() and len(r) > 8:
        try:
            res.append((cur_file, int(r[0]), r[1].strip()[:90], int(r[7]), int(r[6])))
        except ValueError:
            pass
tot_i = sum(x[3] for x in res) or 1
tot_s = sum(x[4] for x in res) or 1
print('total warp-instructions %d, samples %d' % (tot_i, tot_s))
res.sort(key=lambda x: -x[3])
print('%-16s %5s %7s %7s  %s' % ('file', 'line', 'inst%', 'smpl%', 'source'))
for f, ln, src, ins, smp in res[:top]:
    print('%-16s %5d %6.2f%% %6.2f%%  %s' % (f, ln, 100.0 * ins / tot_i, 100.0 * smp / tot_s, src))
