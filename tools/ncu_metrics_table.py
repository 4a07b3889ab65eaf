#!/usr/bin/env python
"""Compact table from an `ncu --csv --metrics ... --log-file X.csv` launch log: one row per kernel (last launch of each
name) with the requested metrics.  usage: tools/ncu_metrics_table.py log.csv [log2.csv ...] > summary.csv"""
import collections
import csv
import sys

SHORT = {
    'gpu__time_duration.sum': 'time_us',
    'smsp__inst_executed.sum': 'warp_inst',
    'sm__inst_executed_pipe_tensor.sum': 'tensor_inst',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_pct',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active': 'fma_pipe_pct',
    'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active': 'fmaheavy_pipe_pct',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active': 'alu_pipe_pct',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active': 'fp64_pipe_pct',
    'sm__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'smem_wavefronts',
}
cols = list(SHORT.values())
print(','.join(['kernel'] + cols))
for path in sys.argv[1:]:
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    last = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split('(')[0].replace('void ', '')
        last.setdefault(name, {})
        last[name].setdefault(r[ii], {})[r[mi]] = r[vi].replace(',', '')
    for name, launches in last.items():
        m = launches[sorted(launches, key=int)[-1]]
        vals = []
        for metric, short in SHORT.items():
            v = m.get(metric, '')
            if short == 'time_us' and v:
                v = '%.1f' % (float(v) / 1e3)
            vals.append(v)
        print(','.join([name] + vals))
