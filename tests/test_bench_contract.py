"""The reference arm of bench.py runs on host cores only: check here (no GPU) that it keeps the driver's
contract - exactly one JSON line on stdout carrying the agreed keys - for the default workload."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '0', *extra], capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run()
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'monte_carlo_realizations_per_s'
    assert d['unit'] == 'realizations/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['config']['workload'] == 'ofdm1024_qam64_mimo2x2_tdl'
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['steps'] == 1 and d['warmup'] == 0 and d['n_gpus'] == 1


def test_reference_arm_non_zero_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without output."""
    out = _run(env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert out.strip() == ''
