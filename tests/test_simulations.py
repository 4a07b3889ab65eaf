"""Host-side Monte Carlo driver (pyphysim_b200.simulations) — CPU tests, modelled on the reference's
tests/simulations_package_test.py (dummy runners, SkipThisOne, partial-result resume, per-index run)
plus BASELINE config C1: 16-QAM over AWGN, 1e5 symbols, one SNR point, through SimulationRunner with
NumPy arithmetic (the oracle stands in for the modulator: no GPU in this test)."""
import math
import os

import numpy as np
import pytest

from pyphysim_b200.simulations import (Result, SimulationParameters, SimulationResults, SimulationRunner,
                                       SkipThisOne, combine_simulation_parameters, counters_to_results,
                                       get_partial_results_filename)


# ------------------------------------------------------------------ Result
def test_result_update_types():
    r = Result("sum", Result.SUMTYPE)
    assert r.get_result() == "Nothing yet"
    r.update(13)
    r.update(4)
    assert r.get_result() == 17 and r.num_updates == 2 and r.type_name == 'SUMTYPE'
    assert r.get_result_mean() == 8.5 and abs(r.get_result_var() - ((13 ** 2 + 4 ** 2) / 2 - 8.5 ** 2)) < 1e-12
    q = Result("ratio", Result.RATIOTYPE)
    q.update(3, 10)
    q.update(6, 7)
    assert q.get_result() == 9 / 17 and abs(q.get_result_mean() - (0.3 + 6 / 7) / 2) < 1e-12
    with pytest.raises(ValueError):
        q.update(3)
    m = Result("misc", Result.MISCTYPE)
    m.update(0.4)
    m.update(0.1)
    assert m.get_result() == 0.1
    with pytest.raises(RuntimeError):
        m.get_confidence_interval()
    c = Result.create("choice", Result.CHOICETYPE, 2, 4)       # np.int bug of the reference avoided
    c.update(2)
    c.update(np.int64(0))
    np.testing.assert_allclose(c.get_result(), [1 / 3, 0, 2 / 3, 0])
    with pytest.raises(RuntimeError):
        Result("choice", Result.CHOICETYPE)
    lo, hi = q.get_confidence_interval(95.0)
    assert lo < q.get_result_mean() < hi
    acc = Result.create("acc", Result.RATIOTYPE, 1, 4, accumulate_values=True)
    acc.update(2, 5)
    assert acc.get_result_accumulated_values() == [1, 2] and acc.get_result_accumulated_totals() == [4, 5]


def test_result_merge_and_serialisation():
    a = Result.create("e", Result.SUMTYPE, 5)
    a.merge(Result.create("e", Result.SUMTYPE, 7))
    assert a.get_result() == 12 and a.num_updates == 2
    r1 = Result.create("r", Result.RATIOTYPE, 1, 10)
    r1.merge(Result.create("r", Result.RATIOTYPE, 3, 10))
    assert r1.get_result() == 0.2 and abs(r1.get_result_mean() - 0.2) < 1e-12
    m = Result.create("m", Result.MISCTYPE, 1.0)
    m.merge(Result.create("m", Result.MISCTYPE, 2.0))
    assert m.get_result() == 2.0
    with pytest.raises(AssertionError):
        a.merge(r1)
    for r in (a, r1, m, Result.create("c", Result.CHOICETYPE, 1, 3)):
        assert Result.from_json(r.to_json()) == r
    assert repr(r1).startswith("Result -> r: 4/20")


# ------------------------------------------------------------------ SimulationParameters
def test_parameters_unpacking():
    p = SimulationParameters.create({'SNR': np.array([0, 5, 10]), 'M': [4, 16], 'NSymbs': 100})
    with pytest.raises(ValueError):
        p.set_unpack_parameter('NSymbs')
    with pytest.raises(ValueError):
        p.set_unpack_parameter('nope')
    assert p.get_num_unpacked_variations() == 1 and p.get_unpacked_params_list() == [p]
    p.set_unpack_parameter('SNR')
    p.set_unpack_parameter('M')
    assert p.unpacked_parameters == ['M', 'SNR'] and p.fixed_parameters == ['NSymbs']
    lst = p.get_unpacked_params_list()
    assert len(lst) == 6 == p.get_num_unpacked_variations()
    assert [(v['M'], v['SNR']) for v in lst] == [(4, 0), (4, 5), (4, 10), (16, 0), (16, 5), (16, 10)]
    assert [v.unpack_index for v in lst] == list(range(6)) and lst[3]['NSymbs'] == 100
    assert lst[2].get_num_unpacked_variations() == 6
    assert list(p.get_pack_indexes({'M': 16})) == [3, 4, 5]
    assert list(p.get_pack_indexes({'SNR': 5})) == [1, 4]
    assert list(p.get_pack_indexes({'SNR': 5, 'M': 4})) == [1]
    q = SimulationParameters.from_json(p.to_json())
    assert q == p and q.unpacked_parameters == ['M', 'SNR']
    q.add('rep_max', 7)
    p.add('rep_max', 9)
    assert q == p                                   # rep_max is ignored by ==
    q['NSymbs'] = 101
    assert q != p
    p.remove('M')
    assert p.unpacked_parameters == ['SNR'] and len(p) == 3
    assert "'SNR*'" in repr(p)


def test_combine_parameters():
    a = SimulationParameters.create({'p1': 10, 'p2': np.array([1, 2, 3])})
    b = SimulationParameters.create({'p1': 10, 'p2': np.array([2, 4, 6])})
    a.set_unpack_parameter('p2')
    b.set_unpack_parameter('p2')
    u = combine_simulation_parameters(a, b)
    assert list(u['p2']) == [1, 2, 3, 4, 6] and u['p1'] == 10 and u.unpacked_parameters == ['p2']
    b['p1'] = 11
    with pytest.raises(RuntimeError):
        combine_simulation_parameters(a, b)


# ------------------------------------------------------------------ SimulationResults
def test_simulation_results_containers(tmp_path):
    sr = SimulationResults()
    sr.add_new_result("a", Result.SUMTYPE, 1)
    sr.add_new_result("b", Result.RATIOTYPE, 1, 4)
    other = SimulationResults()
    other.add_new_result("a", Result.SUMTYPE, 2)
    other.add_new_result("b", Result.RATIOTYPE, 3, 4)
    sr.merge_all_results(other)
    assert sr['a'][-1].get_result() == 3 and sr['b'][-1].get_result() == 0.5
    sr.append_all_results(other)
    assert len(sr['a']) == 2 and sr.get_result_values_list('a') == [3, 2]
    with pytest.raises(ValueError):
        sr.append_result(Result.create("a", Result.RATIOTYPE, 1, 2))
    skipped = SimulationResults()
    skipped.add_new_result("a", Result.SUMTYPE, 0)
    skipped.add_new_result("b", Result.RATIOTYPE, 0, 1)
    skipped.add_new_result("num_skipped_reps", Result.SUMTYPE, 1)
    sr.merge_all_results(skipped)
    assert sr['num_skipped_reps'][-1].get_result() == 1
    p = SimulationParameters.create({'SNR': np.array([0, 5]), 'x': 3})
    p.set_unpack_parameter('SNR')
    sr2 = SimulationResults()
    sr2.set_parameters(p)
    for v in (10, 20):
        sr2.append_result(Result.create("errs", Result.SUMTYPE, v))
    assert sr2.get_result_values_list('errs', {'SNR': 5}) == [20]
    assert len(sr2.get_result_values_confidence_intervals('errs')) == 2
    for ext in ('pickle', 'json'):
        name = sr2.save_to_file(str(tmp_path / ('res_{x}.' + ext)))
        assert name.endswith('res_3.' + ext)
        back = SimulationResults.load_from_file(name)
        assert back == sr2 and back.params == p
    assert sr2.get_filename_with_replaced_params('a_{SNR}') == 'a_[0,5]'
    with pytest.raises(ValueError):
        sr2.set_parameters({'a': 1})
    df = sr2.to_dataframe()
    assert list(df['errs']) == [10, 20] and list(df['SNR']) == [0, 5]
    c = counters_to_results([3, 5, 100, 400])
    assert c['ser'][0].get_result() == 0.03 and c['ber'][0].get_result() == 5 / 400
    assert sorted(c.get_result_names()) == ['ber', 'bit_errors', 'num_bits', 'num_symbols', 'ser',
                                            'symbol_errors']


# ------------------------------------------------------------------ runner
class _DummyRunner(SimulationRunner):
    """Deterministic results, like _DummyRunner in the reference's tests."""

    def __init__(self):
        super().__init__(read_command_line_args=False)
        self.rep_max = 2
        self.update_progress_function_style = None
        self.params.add('SNR', np.array([0., 5., 10., 15., 20.]))
        self.params.add('bias', 1.3)
        self.params.add('extra', np.array([2.2, 4.1]))
        self.params.set_unpack_parameter('SNR')
        self.params.set_unpack_parameter('extra')
        self.calls = 0

    def _run_simulation(self, current_params):
        self.calls += 1
        value = 1.2 * current_params['SNR'] + current_params['bias'] + current_params['extra']
        sr = SimulationResults()
        sr.add_new_result('lala', Result.SUMTYPE, value)
        return sr


def test_runner_serial(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    r = _DummyRunner()
    r.set_results_filename('dummy_{bias}')
    r.partial_results_folder = 'partial'
    r.delete_partial_results_bool = True
    r.simulate()
    assert r.runned_reps == [2] * 10 and r.calls == 20
    expected = [2 * (1.2 * snr + 1.3 + e) for snr in (0., 5., 10., 15., 20.) for e in (2.2, 4.1)]
    np.testing.assert_allclose(r.results.get_result_values_list('lala'), expected)
    assert 'elapsed_time' in r.results.get_result_names() and r.results.params['rep_max'] == 2
    assert all(v == 0 for v in r.results.get_result_values_list('num_skipped_reps'))
    assert os.path.exists(r.results_filename) and r.results_filename.endswith('dummy_1.3.pickle')
    assert not os.listdir(r.partial_results_folder)                 # partial files deleted
    back = SimulationResults.load_from_file(r.results_filename)
    np.testing.assert_allclose(back.get_result_values_list('lala'), expected)
    assert back.runned_reps == [2] * 10 and isinstance(r.elapsed_time, str)
    with pytest.raises(NotImplementedError):
        SimulationRunner(read_command_line_args=False).simulate()


def test_runner_resume_from_partial_results(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    r = _DummyRunner()
    r.set_results_filename('resume')
    r.partial_results_folder = 'partial'
    r.simulate()
    assert len(os.listdir(r.partial_results_folder)) == 10
    r2 = _DummyRunner()                                              # more repetitions: continues
    r2.rep_max = 4
    r2.set_results_filename('resume')
    r2.partial_results_folder = 'partial'
    r2.simulate()
    assert r2.calls == 20 and r2.runned_reps == [4] * 10             # only the 2 extra reps each
    np.testing.assert_allclose(r2.results.get_result_values_list('lala')[0], 4 * (1.3 + 2.2))
    r3 = _DummyRunner()                                              # changed parameters: refuse
    r3.params.add('bias', 1.4)
    r3.set_results_filename('resume')
    r3.partial_results_folder = 'partial'
    with pytest.raises(ValueError):
        r3.simulate()


def test_runner_single_variation_index(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    r = _DummyRunner()
    with pytest.raises(RuntimeError):
        r.simulate(3)                                                # needs a results filename
    r.set_results_filename('idx')
    r.partial_results_folder = 'partial'
    r.simulate(param_variation_index=3)
    assert r.runned_reps == 2 and r.calls == 2
    name = get_partial_results_filename(r.results_base_filename, r.params.get_unpacked_params_list()[3],
                                        r.partial_results_folder)
    assert name.endswith('idx_unpack_03.pickle')
    part = SimulationResults.load_from_file(name)
    assert part.current_rep == 2 and abs(part['lala'][0].get_result() - 2 * (1.2 * 5 + 1.3 + 4.1)) < 1e-12


class _SkipRunner(SimulationRunner):
    def __init__(self):
        super().__init__(read_command_line_args=False)
        self.rep_max = 5
        self.update_progress_function_style = None
        self.n = 0

    def _run_simulation(self, current_params):
        self.n += 1
        if self.n in (2, 5):
            raise SkipThisOne('nope')
        sr = SimulationResults()
        sr.add_new_result('v', Result.SUMTYPE, 1)
        return sr


def test_skip_this_one_and_keep_going():
    r = _SkipRunner()
    r.simulate()
    assert r.runned_reps == [5] and r.n == 7                          # 5 counted + 2 skipped
    assert r.results['v'][0].get_result() == 5 and r.results['num_skipped_reps'][0].get_result() == 2

    class Early(_DummyRunner):
        def _keep_going(self, current_params, current_sim_results, current_rep):
            return current_sim_results['lala'][-1].get_result() < 10
    e = Early()
    e.rep_max = 50
    e.simulate()
    assert e.runned_reps[0] == 3 and e.runned_reps[-1] == 1           # 3.5/rep vs 29.4 at the first rep
    with pytest.raises(RuntimeError):
        e.simulate_in_parallel()
    with pytest.raises(RuntimeError):
        e.wait_parallel_simulation()

    class View:                                                       # minimal ipyparallel-like view
        def map(self, f, *iters, block=False):
            return [f(*a) for a in zip(*iters)]
    d = _DummyRunner()
    d.simulate_in_parallel(View())
    assert d.runned_reps == [2] * 10


# ------------------------------------------------------------------ BASELINE config C1 (CPU plumbing)
def test_c1_qam16_awgn_through_runner():
    from oracle import links as OL
    from oracle import modulators as md
    from oracle import philox

    class QamAwgn(SimulationRunner):
        """apps/awgn_modulators/simulate_psk.py:15-115 with the modulator swapped as simulate_qam.py does."""

        def __init__(self):
            super().__init__(read_command_line_args=False)
            self.modulator = OL.Modem('qam', 16)
            self.NSymbs = 100000
            self.rep_max = 1
            self.update_progress_function_style = None
            self.params.add('SNR', np.array([10.0]))
            self.params.set_unpack_parameter('SNR')

        def _run_simulation(self, current_parameters):
            M, snr = self.modulator.M, current_parameters['SNR']
            data = philox.data_indices(1, [0], self.NSymbs, 4)[0]
            tx = self.modulator.modulate(data)
            noise = philox.cnormal(1, 2, [0], self.NSymbs)[0] * np.sqrt(1.0 / md.dB2Linear(snr))
            hat = self.modulator.demodulate(tx + noise)
            se, be = int(np.sum(data != hat)), int(md.count_bit_errors(data, hat))
            return counters_to_results([se, be, data.size, data.size * md.level2bits(M)])

    r = QamAwgn()
    r.simulate()
    res = r.results
    assert res['num_symbols'][0].get_result() == 100000 and res['num_bits'][0].get_result() == 400000
    ser = res['ser'][0].get_result()
    theory = 1 - (1 - 2 * (1 - 1 / 4) * 0.5 * math.erfc(math.sqrt(10.0 * 3 / 15) / math.sqrt(2))) ** 2
    assert abs(ser - theory) < 0.01 and res['ber'][0].get_result() < ser
    assert res['symbol_errors'][0].get_result() == round(ser * 100000)
