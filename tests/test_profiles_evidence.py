"""The committed evidence under profiles/ stays consistent with what bench.py reads and prints (no GPU needed):
per-kernel ncu JSONs carry the fields bench.py quotes and a kernel name in the launcher's format, and the committed
bench lines satisfy the output contract (keys, units, directions) the driver parses."""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, 'profiles')


def test_ncu_jsons_have_the_fields_bench_quotes():
    files = sorted(glob.glob(os.path.join(PROF, 'ncu_*.json')))
    assert len(files) >= 10
    for f in files:
        d = json.load(open(f))
        for key in ('kernel', 'units_per_launch', 'warp_inst_per_unit', 'dram_bytes_per_unit', 'issue_active_pct',
                    'tensor_pipe_pct', 'registers_per_thread', 'duration_us'):
            assert key in d, (f, key)
        assert re.fullmatch(r'[a-z0-9_]+<[a-z0-9,]+>', d['kernel']), (f, d['kernel'])     # b200phy_last_kernel() format
        assert d['units_per_launch'] > 0 and d['warp_inst_per_unit'] > 0
    tc = json.load(open(os.path.join(PROF, 'ncu_ofdm1024_qam64_mimo2x2_tdl_tcgen05.json')))
    assert tc['kernel'].endswith(',1>') and tc['tensor_pipe_pct'] > 0                      # the tcgen05 variant really used the pipe
    base = json.load(open(os.path.join(PROF, 'ncu_ofdm1024_qam64_mimo2x2_tdl.json')))
    assert base['kernel'].endswith(',0>') and base['tensor_pipe_pct'] == 0
    assert tc['warp_inst_per_unit'] < base['warp_inst_per_unit']


def _lines():
    for f in sorted(glob.glob(os.path.join(PROF, 'bench_full*_r02*.json'))):
        yield f, json.loads(open(f).read().strip().splitlines()[-1])


def test_committed_bench_lines_satisfy_the_contract():
    seen = 0
    for f, d in _lines():
        seen += 1
        for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                    'vs_baseline', 'dtype', 'data', 'config', 'roofline', 'e2e', 'gpu_launches', 'clocks'):
            assert key in d, (f, key)
        assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
        assert d['dtype'] == 'f32' and d['data'] == 'synthetic' and 'workload' in d['config']
        r = d['roofline']
        assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
        e = d['e2e']
        assert e['unit'] == d['unit'] and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
        assert 0.5 * d['fused_rng']['value'] < e['value'] <= 1.05 * d['fused_rng']['value']
        assert d['gpu_launches'] >= d['steps'] and not set(d['clocks']['reasons']) & {
            'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
        if d['n_gpus'] == 1 and 'cpu_baseline' in d:
            c = d['cpu_baseline']
            assert c['kind'] in ('port', 'reference') and c['cores'] >= 1 and c['value'] > 0 and c['sample']
        if 'tensor_core_variant' in d:
            assert d['tensor_core_variant']['adopted'] is False and d['tensor_core_variant']['vs_default_stream'] < 1.05
    assert seen >= 3
