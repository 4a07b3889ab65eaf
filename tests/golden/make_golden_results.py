#!/usr/bin/env python
"""Result files written by the UNMODIFIED reference ("keeps pickle/JSON on-disk formats
loadable by SimulationResults.load_from_file").

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden_results.py

1. builds a SimulationResults with the reference's own classes (an SNR sweep with SUM / RATIO / MISC results,
   several updates per point so the confidence-interval sums are populated), saves it with the reference's
   `save_to_file` as pickle (protocol 2) and JSON  -> tests/golden/results_from_reference.{pickle,json};
2. the REVERSE direction: has pyphysim_b200 write the same content, loads those files with the reference's
   `load_from_file` in a subprocess and records what the reference read back -> results_roundtrip.json.
tests/test_results_files.py loads (1) with pyphysim_b200 and checks (2).
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

STUB = '''
class VdtTypeError(Exception): pass
class VdtValueTooSmallError(Exception): pass
class VdtValueTooBigError(Exception): pass
def is_float(v, *a, **k): return float(v)
def is_integer(v, *a, **k): return int(v)
'''

# what both sides store: three SNR points, three updates each
CONTENT = {
    'SNR': [0.0, 5.0, 10.0], 'M': 16, 'NSymbs': 1000,
    'symbol_errors': [[310, 295, 301], [120, 131, 118], [14, 9, 11]],
    'num_symbols': 1000,
    'runned_reps': [3, 3, 3],
    'label': ['zero', 'five', 'ten'],
}

BUILD = '''
import numpy as np
def build(Result, SimulationParameters, SimulationResults, C):
    params = SimulationParameters.create({'SNR': np.array(C['SNR']), 'M': C['M'], 'NSymbs': C['NSymbs']})
    params.set_unpack_parameter('SNR')
    res = SimulationResults()
    res.set_parameters(params)
    for i in range(len(C['SNR'])):
        se = Result('symbol_errors', Result.SUMTYPE)
        ns = Result('num_symbols', Result.SUMTYPE)
        ser = Result('ser', Result.RATIOTYPE)
        for v in C['symbol_errors'][i]:
            se.update(v); ns.update(C['num_symbols']); ser.update(v, C['num_symbols'])
        res.append_result(se); res.append_result(ns); res.append_result(ser)
        res.append_result(Result.create('label', Result.MISCTYPE, C['label'][i]))
    res.runned_reps = C['runned_reps']
    return res

def summary(res):
    out = {'names': sorted(res.get_result_names()), 'runned_reps': list(res.runned_reps),
           'SNR': [float(v) for v in res.params['SNR']], 'unpacked': sorted(res.params.unpacked_parameters)}
    for n in ('symbol_errors', 'num_symbols', 'ser', 'label'):
        out[n] = [r.get_result() for r in res[n]]
    out['ser_num_updates'] = [r.num_updates for r in res['ser']]
    out['ser_ci'] = [[float(x) for x in r.get_confidence_interval()] for r in res['ser']]
    return out
'''

REF_SIDE = '''
import json, sys
sys.dont_write_bytecode = True
sys.path.insert(0, sys.argv[1])          # stub `validate`
sys.path.insert(0, '/root/reference')
from pyphysim.simulations.results import Result, SimulationResults
from pyphysim.simulations.parameters import SimulationParameters
exec(open(sys.argv[2]).read())
C = json.loads(sys.argv[3])
mode, out = sys.argv[4], sys.argv[5]
if mode == 'write':
    res = build(Result, SimulationParameters, SimulationResults, C)
    res.save_to_file(out + '.pickle')
    res.save_to_file(out + '.json')
    print(json.dumps(summary(res)))
else:
    print(json.dumps({ext: summary(SimulationResults.load_from_file(out + '.' + ext)) for ext in ('pickle', 'json')}))
'''


def main():
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, 'validate.py'), 'w').write(STUB)
        open(os.path.join(tmp, 'build.py'), 'w').write(BUILD)
        open(os.path.join(tmp, 'ref_side.py'), 'w').write(REF_SIDE)

        def ref(mode, out):
            r = subprocess.run([sys.executable, os.path.join(tmp, 'ref_side.py'), tmp, os.path.join(tmp, 'build.py'),
                                json.dumps(CONTENT), mode, out], capture_output=True, text=True, cwd=tmp)
            if r.returncode:
                raise RuntimeError(r.stderr)
            return json.loads(r.stdout.strip().splitlines()[-1])

        # 1. the reference writes
        ref_summary = ref('write', os.path.join(HERE, 'results_from_reference'))
        # 2. pyphysim_b200 writes, the reference reads
        sys.path.insert(0, ROOT)
        from pyphysim_b200.simulations import Result, SimulationParameters, SimulationResults
        ns = {}
        exec(BUILD, ns)
        ours = ns['build'](Result, SimulationParameters, SimulationResults, CONTENT)
        base = os.path.join(tmp, 'ours')
        ours.save_to_file(base + '.json')
        # save_to_file writes the reference's class paths into the pickle (simulations/results.py:
        # _ReferencePathPickler), which is what a reference-side loader (bin/combine_results.py) resolves
        ours.save_to_file(base + '.pickle')
        back = ref('read', base)
        json.dump({'content': CONTENT, 'reference_summary': ref_summary,
                   'reference_read_of_b200_files': back,
                   },
                  open(os.path.join(HERE, 'results_roundtrip.json'), 'w'), indent=1, sort_keys=True)
    print('wrote results_from_reference.{pickle,json}, results_roundtrip.json')


if __name__ == '__main__':
    main()
