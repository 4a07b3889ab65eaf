"""Result files cross the boundary in both directions (reference
pyphysim/simulations/results.py:1454-1596): files written by the UNMODIFIED reference
(tests/golden/results_from_reference.{pickle,json}, made by tests/golden/make_golden_results.py) load into
pyphysim_b200, and files written by pyphysim_b200 were loaded by the reference when the fixture was made
(tests/golden/results_roundtrip.json records what the reference read back)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

from pyphysim_b200.simulations import Result, SimulationParameters, SimulationResults


def _summary(res):
    out = {'names': sorted(res.get_result_names()), 'runned_reps': list(res.runned_reps),
           'SNR': [float(v) for v in res.params['SNR']], 'unpacked': sorted(res.params.unpacked_parameters)}
    for n in ('symbol_errors', 'num_symbols', 'ser', 'label'):
        out[n] = [r.get_result() for r in res[n]]
    out['ser_num_updates'] = [r.num_updates for r in res['ser']]
    out['ser_ci'] = [[float(x) for x in r.get_confidence_interval()] for r in res['ser']]
    return out


def _build(C):
    params = SimulationParameters.create({'SNR': np.array(C['SNR']), 'M': C['M'], 'NSymbs': C['NSymbs']})
    params.set_unpack_parameter('SNR')
    res = SimulationResults()
    res.set_parameters(params)
    for i in range(len(C['SNR'])):
        se, ns, ser = (Result('symbol_errors', Result.SUMTYPE), Result('num_symbols', Result.SUMTYPE),
                       Result('ser', Result.RATIOTYPE))
        for v in C['symbol_errors'][i]:
            se.update(v)
            ns.update(C['num_symbols'])
            ser.update(v, C['num_symbols'])
        for r in (se, ns, ser, Result.create('label', Result.MISCTYPE, C['label'][i])):
            res.append_result(r)
    res.runned_reps = C['runned_reps']
    return res


@pytest.fixture(scope='module')
def roundtrip():
    return json.load(open(os.path.join(GOLDEN, 'results_roundtrip.json')))


@pytest.mark.parametrize('ext', ['pickle', 'json'])
def test_files_written_by_the_reference_load_here(ext, roundtrip):
    res = SimulationResults.load_from_file(os.path.join(GOLDEN, 'results_from_reference.' + ext))
    got, ref = _summary(res), roundtrip['reference_summary']
    for k in ref:
        if k == 'ser_ci' or k == 'ser':
            np.testing.assert_allclose(got[k], ref[k], rtol=1e-12)
        else:
            assert got[k] == ref[k], k
    # a loaded result keeps working: merge another update into it and re-save
    res['ser'][0].update(10, 1000)
    assert res['ser'][0].num_updates == 4 and res['ser'][0].get_result() == (906 + 10) / 4000


def test_files_written_here_were_loaded_by_the_reference(roundtrip, tmp_path):
    """The reverse direction is executed by make_golden_results.py (the reference cannot travel); here: what it
    recorded equals the content, and today's writer still produces the same bytes-level format."""
    ref = roundtrip['reference_summary']
    for ext in ('pickle', 'json'):
        assert roundtrip['reference_read_of_b200_files'][ext] == ref
    ours = _build(roundtrip['content'])
    assert _summary(ours)['symbol_errors'] == ref['symbol_errors']
    # pickle: protocol 2, the reference's class paths, none of ours
    p = ours.save_to_file(str(tmp_path / 'ours.pickle'))
    raw = open(p, 'rb').read()
    assert raw[:2] == b'\x80\x02'
    assert b'pyphysim.simulations.results' in raw and b'pyphysim.simulations.parameters' in raw
    assert b'pyphysim_b200' not in raw
    assert _summary(SimulationResults.load_from_file(p)) == _summary(ours)
    # json: same schema as the file the reference wrote (keys, array / set encodings)
    j = json.load(open(ours.save_to_file(str(tmp_path / 'ours.json'))))
    r = json.load(open(os.path.join(GOLDEN, 'results_from_reference.json')))
    assert j['params'] == r['params'] and j['runned_reps'] == r['runned_reps']
    assert set(r) <= set(j) and j['results'].keys() == r['results'].keys()
    for name in r['results']:
        for a, b in zip(j['results'][name], r['results'][name]):
            assert a.keys() == b.keys()
            for k in b:
                if isinstance(b[k], float):
                    assert a[k] == pytest.approx(b[k], rel=1e-12)
                else:
                    assert a[k] == b[k], (name, k)


def test_partial_result_files_use_the_same_format(tmp_path):
    """Partial results (runner.py:996-1069) go through the same writer, so a reference-side
    bin/combine_results.py can merge them."""
    from pyphysim_b200.simulations import SimulationRunner, counters_to_results

    class R(SimulationRunner):
        def __init__(self):
            super().__init__(read_command_line_args=False)
            self.rep_max = 2
            self.params.add('SNR', np.array([0., 10.]))
            self.params.set_unpack_parameter('SNR')
            self.update_progress_function_style = None
            self.partial_results_folder = str(tmp_path / 'partial')

        def _run_simulation(self, p):
            return counters_to_results([5, 7, 100, 400])

    r = R()
    r.set_results_filename(str(tmp_path / 'final'))
    r.simulate()
    # an absolute results filename wins over the folder in os.path.join, exactly as in the reference
    # (runner.py:109-145): the partial files sit next to the final one
    files = sorted(f for f in os.listdir(tmp_path) if '_unpack_' in f)
    assert files == ['final_unpack_0.pickle', 'final_unpack_1.pickle']
    raw = open(os.path.join(tmp_path, files[0]), 'rb').read()
    assert b'pyphysim.simulations.results' in raw and b'pyphysim_b200' not in raw
    back = SimulationResults.load_from_file(os.path.join(tmp_path, files[1]))
    assert back.current_rep == 2 and back['symbol_errors'][0].get_result() == 10


@pytest.mark.skipif(not os.path.isdir('/root/reference/pyphysim'), reason='needs the reference checkout (build container only)')
def test_reverse_direction_live(tmp_path, roundtrip):
    """Build container only: the unmodified reference loads files written right now."""
    import subprocess
    import sys
    ours = _build(roundtrip['content'])
    base = str(tmp_path / 'ours')
    ours.save_to_file(base + '.pickle')
    ours.save_to_file(base + '.json')
    (tmp_path / 'validate.py').write_text(
        'class VdtTypeError(Exception): pass\nclass VdtValueTooSmallError(Exception): pass\n'
        'class VdtValueTooBigError(Exception): pass\ndef is_float(v, *a, **k): return float(v)\n'
        'def is_integer(v, *a, **k): return int(v)\n')
    code = ("import sys, json; sys.dont_write_bytecode = True; sys.path.insert(0, %r); sys.path.insert(0, '/root/reference');"
            "from pyphysim.simulations.results import SimulationResults as S;"
            "print(json.dumps([[r.get_result() for r in S.load_from_file(%r + e)['symbol_errors']] for e in ('.pickle', '.json')]))"
            % (str(tmp_path), base))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr
    assert json.loads(out.stdout.strip().splitlines()[-1]) == [[906, 369, 34]] * 2
