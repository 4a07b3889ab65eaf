#!/usr/bin/env python
"""Static view of the hot loops of a kernel: for every backward branch in the SASS of `function` (substring
match) in an object file, the opcode mix of the loop body.  usage: tools/sass_loops.py file.o substring [min_ffma2]"""
import collections
import re
import subprocess
import sys

obj, sub = sys.argv[1], sys.argv[2]
min_f = int(sys.argv[3]) if len(sys.argv) > 3 else 32
out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
fn, ins = None, []
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
    if m and fn and sub in fn:
        ins.append((int(m.group(1), 16), m.group(2)))
addr = {a: i for i, (a, _) in enumerate(ins)}
print('%d instructions' % len(ins))
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)', t)
    if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
        j = addr[int(m.group(1), 16)]
        c = collections.Counter()
        for _, tt in ins[j:i + 1]:
            tt = re.sub(r'^@!?U?P\d+\s+', '', tt)
            c[tt.split()[0].split('.')[0]] += 1
        if c['FFMA2'] + c['FFMA'] + c['FADD2'] >= min_f:
            print('loop %d..%d (%d instr): %s' % (j, i, i - j + 1, ', '.join('%s %d' % kv for kv in c.most_common(12))))
