#!/usr/bin/env python
"""Per BAR-delimited region: share of executed instructions vs share of warp-stall samples (time)."""
import csv, subprocess, sys
rep = sys.argv[1]
which = sys.argv[2] if len(sys.argv) > 2 else '0'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', which, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(out.splitlines())]
hdr = next(r for r in rows if r and r[0] == 'Address')
k = []
seen_hdr = 0
for r in rows:                                  # the page lists each kernel once per source view: keep the first listing
    if r and r[0] == 'Address':
        seen_hdr += 1
    elif r and r[0].startswith('0x') and seen_hdr == 1:
        k.append(r)
i_ins, i_smp = hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
ti = sum(int(r[i_ins]) for r in k); ts = sum(int(r[i_smp]) for r in k)
seg_i = seg_s = 0; n0 = 0; st = [0] * len(stall_cols)
print('%-14s %7s %7s  top stall reasons (share of region samples)' % ('sass range', 'inst%', 'time%'))
for n, r in enumerate(k):
    seg_i += int(r[i_ins]); seg_s += int(r[i_smp])
    for j, (ci, _) in enumerate(stall_cols):
        st[j] += int(r[ci])
    txt = r[1].strip(); op = (txt.split()[1] if txt.startswith('@') else txt.split()[0]).split('.')[0]
    if op == 'BAR' or n == len(k) - 1:
        if seg_s > 0.005 * ts:
            tot = sum(st) or 1
            top = sorted(zip(st, [h for _, h in stall_cols]), reverse=True)[:4]
            print('%6d-%6d %6.2f%% %6.2f%%  %s' % (n0, n, 100.0 * seg_i / ti, 100.0 * seg_s / ts,
                  ', '.join('%s %.0f%%' % (h.replace('stall_', ''), 100.0 * v / tot) for v, h in top)))
        seg_i = seg_s = 0; n0 = n + 1; st = [0] * len(stall_cols)
