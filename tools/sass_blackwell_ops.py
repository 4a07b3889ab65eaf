#!/usr/bin/env python
"""Per kernel of libb200phy.so: counts of the SASS mnemonics that prove the Blackwell-specific paths.
usage: cuobjdump -sass pyphysim_b200/libb200phy.so | python tools/sass_blackwell_ops.py > profiles/sass_blackwell_ops_rNN.txt"""
import collections
import re
import subprocess
import sys

OPS = ('UBLKCP', 'UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTCATOMSWS', 'SYNCS', 'FFMA2', 'FADD2', 'FMUL2', 'LDGSTS', 'DFMA')
cur, cnt, tot = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        tot[cur] += 1
        for o in OPS:
            if m.group(1).startswith(o):
                cnt[cur][o] += 1
print('# SASS mnemonics that prove the Blackwell-specific paths, per kernel of pyphysim_b200/libb200phy.so (cuobjdump -sass)')
print('# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st,')
print('# UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc / dealloc, FFMA2 / FADD2 / FMUL2 = packed f32x2 arithmetic, LDGSTS = cp.async')
names = [k for k in sorted(tot, key=lambda k: -tot[k]) if any(cnt[k][o] for o in ('UBLKCP', 'UTCHMMA', 'FFMA2', 'LDTM'))]
dem = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.splitlines()
for k, d in zip(names, dem):
    d = re.sub(r'\(.*', '', d).replace('b200phy::', '').replace('void ', '').replace('(bool)', '').replace('(int)', '')
    print('%-58s %5d instr | %s' % (d[:58], tot[k], ', '.join('%s %d' % (o, cnt[k][o]) for o in OPS if cnt[k][o])))
