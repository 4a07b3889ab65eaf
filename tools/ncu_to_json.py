#!/usr/bin/env python
"""profiles/ncu_<workload>.json from an `ncu --set full` capture: the per-unit instruction and DRAM figures bench.py
quotes in `roofline.traffic` / `issue`, keyed by the kernel instantiation they were measured on.

    python tools/ncu_to_json.py report.ncu-rep UNITS_PER_LAUNCH [kernel-regex | #N] > profiles/ncu_<workload>.json

bench.py compares the `kernel` field with b200phy_last_kernel() of its own launch and refuses a capture of a
different instantiation, so a profile that went stale with a kernel change is never quoted."""
import csv
import json
import re
import subprocess
import sys


def normalise(name):
    """'void b200phy::ofdm_tdl_pair_kernel<(bool)0, (int)2, ...>(args)' -> 'ofdm_tdl_pair_kernel<0,2,...>'"""
    name = re.sub(r'^void\s+', '', name.strip())
    depth, end = 0, len(name)
    for i, ch in enumerate(name):                    # cut the argument list: first '(' at template depth 0
        if ch == '<':
            depth += 1
        elif ch == '>':
            depth -= 1
        elif ch == '(' and depth == 0:
            end = i
            break
    name = name[:end]
    name = name.split('::')[-1] if '<' not in name else name[name.rfind('::', 0, name.index('<')) + 2:] \
        if '::' in name[:name.index('<')] else name
    name = re.sub(r'\((bool|int|unsigned int|long)\)', '', name)
    name = name.replace('true', '1').replace('false', '0')
    return re.sub(r'\s+', '', name)


def main():
    rep, units = sys.argv[1], int(sys.argv[2])
    sel = sys.argv[3] if len(sys.argv) > 3 else None          # kernel-name regex, or '#N' = N-th captured launch
    rx = re.compile(sel) if sel and not sel.startswith('#') else None
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    if rx:
        body = [r for r in body if rx.search(r[hdr.index('Kernel Name')])]
    r = body[int(sel[1:])] if sel and sel.startswith('#') else body[-1]

    def g(metric):
        return float(r[hdr.index(metric)].replace(',', '')) if metric in hdr else None

    def to_bytes(metric):
        v, u = g(metric), rows[1][hdr.index(metric)].split('/')[0]
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)

    dur, du = g('gpu__time_duration.sum'), rows[1][hdr.index('gpu__time_duration.sum')]
    dur_us = dur * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'second': 1e6}.get(du, 1)
    rd, wr = to_bytes('dram__bytes_read.sum'), to_bytes('dram__bytes_write.sum')
    d = {"kernel": normalise(r[hdr.index('Kernel Name')]),
         "capture": "ncu --set full --clock-control none (%s)" % rep.split('/')[-1],
         "units_per_launch": units,
         "warp_inst_per_unit": g('smsp__inst_executed.sum') / units,
         "dram_bytes_per_unit": (rd + wr) / units, "dram_read_bytes_per_unit": rd / units,
         "dram_write_bytes_per_unit": wr / units,
         "issue_active_pct": g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
         "fma_pipe_pct": g('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
         "fp64_pipe_pct": g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
         "tensor_pipe_pct": g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
         "smem_pipe_pct": g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
         "dram_pct_of_peak": g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
         "registers_per_thread": g('launch__registers_per_thread'),
         "dyn_smem_bytes": to_bytes('launch__shared_mem_per_block_dynamic'),
         "warps_active_pct": g('sm__warps_active.avg.pct_of_peak_sustained_active'),
         "duration_us": dur_us}
    print(json.dumps(d, indent=1))


if __name__ == '__main__':
    if len(sys.argv) == 3 and sys.argv[1] == '--normalise':
        print(normalise(sys.argv[2]))
    else:
        main()
