"""The pyphysim-API façade classes (modulators / channels / mimo / util) against the golden fixtures
made from the unmodified reference and against the oracle — written the way the reference's own unit
tests are (tests/modulators_package_test.py, channels_package_test.py, mimo_package_test.py)."""
import math

import numpy as np
import pytest

from oracle import fading as ofading
from oracle import philox

pytestmark = pytest.mark.gpu
SEED = 0xC0FFEE
TOL = dict(rtol=1e-10, atol=1e-11)


class QueueRS:
    """RandomState stand-in serving queued arrays (same trick as tests/golden/make_golden.py)."""

    def __init__(self):
        self.queue = []

    def rand(self, *shape):
        n = int(np.prod(shape))
        if self.queue and self.queue[0].size == n:
            return self.queue.pop(0).reshape(shape)
        return np.full(shape, 0.25)


# ------------------------------------------------------------------ modulators
def test_constellations_and_roundtrip(golden):
    from pyphysim_b200.modulators import BPSK, PSK, QAM, QPSK
    g = golden('constellations')
    for M in (4, 16, 64, 256):
        q = QAM(M)
        np.testing.assert_allclose(q.symbols, g['qam%d' % M], **TOL)
        assert q.M == M and abs(q.K - math.log2(M)) < 1e-12 and q.name == '%d-QAM' % M
        idx = np.arange(M)
        mod = q.modulate(idx)
        assert mod.dtype == np.complex128
        np.testing.assert_allclose(mod, g['qam%d' % M], **TOL)
        assert np.array_equal(q.demodulate(mod + 1e-3), idx) and q.demodulate(mod).dtype == np.int64
    for M in (2, 4, 8, 16):
        np.testing.assert_allclose(PSK(M).symbols, g['psk%d' % M], **TOL)
    np.testing.assert_allclose(PSK(8, 0.3).symbols, g['psk8_off'], **TOL)
    p = PSK(8)
    p.setPhaseOffset(0.2)
    np.testing.assert_allclose(p.symbols, g['psk8_setoffset'], **TOL)
    assert np.array_equal(p.demodulate(p.modulate(np.arange(8))), np.arange(8))   # table re-uploaded
    np.testing.assert_allclose(QPSK().symbols, g['qpsk'], **TOL)
    assert list(BPSK().symbols) == [1, -1] and BPSK().name == 'BPSK'
    assert list(BPSK().modulate(np.array([0, 1, 1, 0]))) == [1, -1, -1, 1]
    with pytest.raises(ValueError):
        QAM(32)
    with pytest.raises(ValueError):
        QAM(16).modulate(np.array([0, 16]))
    with pytest.raises(ValueError):
        BPSK().modulate(np.array([0, 2]))
    assert QAM(16).modulate(3) == QAM(16).symbols[3]


def test_demodulate_matches_reference(golden):
    from pyphysim_b200.modulators import BPSK, PSK, QAM, QPSK
    from pyphysim_b200.util import misc
    g = golden('demap')
    mods = {'qam16': QAM(16), 'qam64': QAM(64), 'qam256': QAM(256), 'psk8': PSK(8), 'qpsk': QPSK(),
            'bpsk': BPSK()}
    for name, m in mods.items():
        hat = m.demodulate(g[name + '_r'])
        assert np.array_equal(hat, g[name + '_hat']), name
        assert misc.count_bit_errors(g[name + '_idx'], hat) == int(g[name + '_biterr'])
        se, be = misc.count_symbol_and_bit_errors(g[name + '_idx'], hat)
        assert se == int(np.sum(g[name + '_idx'] != hat)) and be == int(g[name + '_biterr'])
    assert np.array_equal(misc.count_bits(g['count_bits_in']), g['count_bits_out'])
    a = np.arange(24).reshape(4, 6)
    b = a[::-1].copy()
    assert np.array_equal(misc.count_bit_errors(a, b, axis=1),
                          np.array([sum(bin(x ^ y).count('1') for x, y in zip(r1, r2)) for r1, r2 in zip(a, b)]))
    assert misc.level2bits(64) == 6 and misc.int2bits(0) == 1 and misc.count_bits(7) == 3
    import torch
    r = torch.from_numpy(g['qam64_r']).cuda()
    hat = mods['qam64'].demodulate(r)
    assert hat.is_cuda and np.array_equal(hat.cpu().numpy(), g['qam64_hat'])
    hat32 = mods['qam64'].demodulate(g['qam64_r'].astype(np.complex64))
    assert np.mean(hat32 != g['qam64_hat']) < 1e-3


def test_qpsk_slicer_and_table_fallback():
    """QPSK() demaps with the quadrant slicer (B200PHY_MODEM_QPSK): same indices as the oracle's min-distance
    search on 1e6 noisy points; after setPhaseOffset (the reference rebuilds WITHOUT the Gray order,
    fundamental.py:450-459) the object falls back to the table search and still matches."""
    from oracle import modulators as omd
    from pyphysim_b200 import _lib
    from pyphysim_b200.modulators import QPSK
    q = QPSK()
    assert q._kind == _lib.MODEM_QPSK
    r = 1.2 * philox.cnormal(SEED, 2, [77], 1000000)[0]
    ref = omd.demodulate(q.symbols, r)
    assert np.array_equal(q.demodulate(r), ref)
    assert np.array_equal(q.demodulate(q.modulate(np.arange(4))), np.arange(4))
    q.setPhaseOffset(0.3)
    assert q._kind == _lib.MODEM_TABLE
    assert np.array_equal(q.demodulate(r[:100000]), omd.demodulate(q.symbols, r[:100000]))


def test_theoretical_curves():
    from pyphysim_b200.modulators import BPSK, PSK, QAM
    # numbers asserted by the reference's tests (tests/modulators_package_test.py:72-138, 282-322)
    q = QAM(16)
    snr = np.array([0., 5, 10, 15, 20])
    ser = q.calcTheoreticalSER(snr)
    assert np.all(np.diff(ser) < 0) and 0.7 < ser[0] < 0.8 and ser[-1] < 2e-5
    np.testing.assert_allclose(q.calcTheoreticalBER(snr) * 4 / 2, q._calcTheoreticalSingleCarrierErrorRate(snr))
    assert abs(BPSK().calcTheoreticalSER(0.0) - 0.0786496) < 1e-6
    assert abs(PSK(4).calcTheoreticalSER(10.0) - 2 * 0.5 * math.erfc(math.sqrt(10.0) * math.sin(math.pi / 4))) < 1e-12
    assert q.calcTheoreticalSpectralEfficiency(30.0) == q.K


def test_randn_c_statistics():
    from pyphysim_b200.util.misc import randn_c
    x = randn_c(200000)
    assert x.dtype == np.complex128 and x.shape == (200000,)
    assert abs(np.mean(np.abs(x) ** 2) - 1) < 0.01 and abs(np.mean(x)) < 0.01
    assert randn_c(3, 4).shape == (3, 4) and isinstance(randn_c(), complex)
    assert not np.array_equal(randn_c(8), randn_c(8))


# ------------------------------------------------------------------ OFDM
def test_ofdm_matches_reference(golden):
    from pyphysim_b200.modulators import OFDM
    g = golden('ofdm')
    for tag in 'abcd':
        f, c, u = (int(v) for v in g[tag + '_params'])
        o = OFDM(f, c, u)
        assert np.array_equal(o.get_used_subcarrier_indexes(), g[tag + '_bins'])
        np.testing.assert_allclose(o.modulate(g[tag + '_x']), g[tag + '_mod'], **TOL)
        r = g[tag + '_r'].copy()
        np.testing.assert_allclose(o.demodulate(r), g[tag + '_demod'], **TOL)
        assert r.shape == (3, f + c)                      # in-place reshape quirk of the reference
    o = OFDM(64, 16, 52)
    grid = o._prepare_input_signal(np.r_[1:53])
    assert np.array_equal(grid[0], np.r_[0, 27:53, np.zeros(11), 1:27])
    assert o._calc_zeropad(52) == (0, 1) and OFDM(64, 16, 60)._calc_zeropad(52) == (8, 1)
    assert o._calculate_power_scale() == 64.0 ** 2 / 68
    for bad in ((64, 65, 52), (64, 16, 66), (64, 16, 51), (64, -1, 52)):
        with pytest.raises(ValueError):
            OFDM(*bad)


# ------------------------------------------------------------------ fading generators / TDL
def test_profiles(golden):
    from pyphysim_b200.channels import fading
    g = golden('fading')
    for pn, prof in (('tu', fading.COST259_TUx), ('ra', fading.COST259_RAx), ('ht', fading.COST259_HTx)):
        for tn, Ts in (('2048', 1 / (15e3 * 2048)), ('1024', 1 / (15e3 * 1024)), ('128', 1 / (15e3 * 128))):
            d = prof.get_discretize_profile(Ts)
            assert np.array_equal(d.tap_delays, g['%s_%s_delays' % (pn, tn)])
            np.testing.assert_allclose(d.tap_powers_linear, g['%s_%s_powers' % (pn, tn)], rtol=1e-14)
    tu = fading.COST259_TUx.get_discretize_profile(3.255e-08)
    assert list(tu.tap_delays) == [0, 7, 16, 21, 27, 38, 40, 41, 47, 50, 56, 58, 60, 63, 66]
    assert tu.num_taps == 15 and tu.num_taps_with_padding == 67 and tu.is_discretized
    assert tu.name == 'COST259_TU (discretized)'
    with pytest.raises(RuntimeError):
        tu.get_discretize_profile(3.255e-08)
    with pytest.raises(RuntimeError):
        _ = fading.COST259_TUx.num_taps_with_padding
    assert abs(fading.COST259_TUx.rms_delay_spread - 5.000561653134637e-07) < 1e-12


def test_jakes_generator_matches_reference(golden):
    from pyphysim_b200.channels.fading_generators import JakesSampleGenerator, RayleighSampleGenerator
    g = golden('fading')
    rs = QueueRS()
    rs.queue = [g['j0_phi'].reshape(8, 1) / (2 * np.pi), g['j0_psi'].reshape(8, 1) / (2 * np.pi)]
    gen = JakesSampleGenerator(Fd=100.0, Ts=1e-3, L=8, RS=rs)
    assert gen.shape is None and gen.L == 8 and gen.Ts == 1e-3 and gen.Fd == 100.0
    assert gen.get_samples().shape == (1,)
    gen.generate_more_samples(50)
    np.testing.assert_allclose(gen.get_samples(), g['j0_h1'].reshape(-1), **TOL)
    gen.skip_samples_for_next_generation(7)
    gen.generate_more_samples(20)
    np.testing.assert_allclose(gen.get_samples(), g['j0_h2'].reshape(-1), **TOL)
    assert abs(gen._current_time - float(g['j0_t_end'])) < 1e-12
    rs = QueueRS()
    gen = JakesSampleGenerator(Fd=30.0, Ts=5e-6, L=20, RS=rs)
    rs.queue = [g['j1_phi'][..., None] / (2 * np.pi), g['j1_psi'][..., None] / (2 * np.pi)]
    gen.shape = (3, 2)
    gen.generate_more_samples(40)
    np.testing.assert_allclose(gen.get_samples(), g['j1_h'], **TOL)
    # clock: 1 + 100 samples -> 101 Ts (tests/channels_package_test.py:244-259)
    gen = JakesSampleGenerator(Fd=5.0, Ts=1e-3, L=4)
    gen.generate_more_samples(100)
    assert abs(gen._current_time - 101e-3) < 1e-9
    assert gen.get_similar_fading_generator().L == 4
    ray = RayleighSampleGenerator(shape=(3, 2))
    ray.generate_more_samples(1000)
    s = ray.get_samples()
    assert s.shape == (3, 2, 1000) and abs(np.mean(np.abs(s) ** 2) - 1) < 0.1
    assert isinstance(RayleighSampleGenerator().get_samples(), complex)     # randn_c() scalar


def test_tdl_channel_siso_chain_matches_reference(golden):
    from pyphysim_b200.channels import fading
    from pyphysim_b200.channels.fading_generators import JakesSampleGenerator
    from pyphysim_b200.modulators import OFDM, QAM
    from pyphysim_b200.modulators.ofdm import OfdmOneTapEqualizer
    g = golden('tdl')
    fft, cp, used, nsym = (int(v) for v in g['s_params'])
    Ts, Fd = float(g['s_Ts']), float(g['s_Fd'])
    qam, o = QAM(64), OFDM(fft, cp, used)
    tx = o.modulate(qam.modulate(g['s_idx']))
    np.testing.assert_allclose(tx, g['s_tx'], **TOL)
    rs = QueueRS()
    jakes = JakesSampleGenerator(Fd=Fd, Ts=Ts, L=20, RS=rs)
    prof = fading.COST259_TUx.get_discretize_profile(Ts)
    rs.queue = [g['s_phi'][..., None] / (2 * np.pi), g['s_psi'][..., None] / (2 * np.pi)]
    ch = fading.TdlChannel(jakes, prof)
    assert ch.num_taps == prof.num_taps and ch.num_tx_antennas == -1
    with pytest.raises(RuntimeError):
        ch.get_last_impulse_response()
    rx = ch.corrupt_data(tx)
    np.testing.assert_allclose(rx, g['s_rx'], **TOL)
    ir = ch.get_last_impulse_response()
    np.testing.assert_allclose(ir.tap_values_sparse, g['s_taps'], **TOL)
    assert ir.num_samples == tx.size and ir.tap_values.shape == (prof.num_taps_with_padding, tx.size)
    rxn = rx + math.sqrt(float(g['s_nv'])) * g['s_noise']
    Y = o.demodulate(rxn[:tx.size].copy())
    np.testing.assert_allclose(Y, g['s_Y'], **TOL)
    eq = OfdmOneTapEqualizer(o).equalize_data(Y, ir)
    np.testing.assert_allclose(eq, g['s_eq'], rtol=1e-9, atol=1e-10)
    assert np.array_equal(qam.demodulate(eq), g['s_hat'])
    # get_freq_response == np.fft.fft(dense taps) (tests/channels_package_test.py:572-593)
    np.testing.assert_allclose(ir.get_freq_response(fft)[:, ::37],
                               np.fft.fft(ir.tap_values, fft, axis=0)[:, ::37], rtol=1e-9, atol=1e-11)
    with pytest.raises(RuntimeError):
        fading.TdlChannel(JakesSampleGenerator(Ts=1e-3), prof, Ts=2e-3)
    with pytest.raises(TypeError):
        ch.switched_direction = 1


def test_tdl_mimo_channel_matches_reference(golden):
    from pyphysim_b200.channels import fading
    from pyphysim_b200.channels.fading_generators import JakesSampleGenerator
    g = golden('tdl')
    Ts, Fd = float(g['m_Ts']), float(g['m_Fd'])
    rs = QueueRS()
    jakes = JakesSampleGenerator(Fd=Fd, Ts=Ts, L=16, shape=(3, 2), RS=rs)
    prof = fading.COST259_RAx.get_discretize_profile(Ts)
    rs.queue = [g['m_phi'][..., None] / (2 * np.pi), g['m_psi'][..., None] / (2 * np.pi)]
    ch = fading.TdlMimoChannel(jakes, prof)
    assert (ch.num_rx_antennas, ch.num_tx_antennas) == (3, 2)
    y = ch.corrupt_data(g['m_x'])
    np.testing.assert_allclose(y, g['m_y'], **TOL)
    ir = ch.get_last_impulse_response()
    np.testing.assert_allclose(ir.tap_values_sparse, g['m_taps'], **TOL)
    np.testing.assert_allclose(ir.get_freq_response(64)[..., ::50], g['m_freq'], rtol=1e-9, atol=1e-11)
    # explicit shifted multiply-add restatement of the switched direction (fading.py:1098-1106)
    ch.switched_direction = True
    x3 = philox.cnormal(SEED, 1, [77], 3 * 40)[0].reshape(3, 40)
    y2 = ch.corrupt_data(x3)
    taps = ch.get_last_impulse_response().tap_values_sparse
    exp = np.zeros((2, 40 + int(prof.tap_delays[-1])), dtype=complex)
    for i, d in enumerate(prof.tap_delays):
        for r in range(3):
            exp[:, d:d + 40] += taps[i, r, :, :] * x3[r]
    np.testing.assert_allclose(y2, exp, **TOL)
    with pytest.raises(RuntimeError):
        fading.TdlMimoChannel(JakesSampleGenerator(Ts=Ts), prof)


def test_su_channel():
    from pyphysim_b200.channels.singleuser import SuChannel, SuMimoChannel
    from pyphysim_b200.channels.fading_generators import JakesSampleGenerator
    from pyphysim_b200.channels import fading
    su = SuChannel()                                        # flat Rayleigh, Ts = 1
    x = np.ones(16, dtype=complex)
    y = su.corrupt_data(x)
    h = su.get_last_impulse_response().tap_values_sparse
    np.testing.assert_allclose(y, h[0] * x, **TOL)
    su.set_pathloss(0.25)
    y = su.corrupt_data(x)
    h = su.get_last_impulse_response().tap_values_sparse     # includes sqrt(pathloss)
    np.testing.assert_allclose(y, h[0] * x, **TOL)
    assert np.mean(np.abs(h) ** 2) < 2.0
    with pytest.raises(ValueError):
        su.set_pathloss(1.5)
    Ts = 3.255e-8
    su = SuChannel(JakesSampleGenerator(Fd=5, Ts=Ts, L=20), fading.COST259_TUx)
    assert su.num_taps == 15 and su.num_taps_with_padding == 67
    assert su.corrupt_data(np.ones(100, dtype=complex)).shape == (166,)
    mm = SuMimoChannel(2)
    assert mm.corrupt_data(np.ones((2, 10), dtype=complex)).shape == (2, 10)


# ------------------------------------------------------------------ MIMO
def test_blast_and_alamouti_match_reference(golden):
    from pyphysim_b200.mimo import Alamouti, Blast
    g = golden('mimo')
    for name in ('h43', 'h44', 'h22'):
        b = Blast(g[name])
        assert b.getNumberOfLayers() == g[name].shape[1]
        np.testing.assert_allclose(b.encode(g[name + '_s']), g[name + '_x'], **TOL)
        np.testing.assert_allclose(b.decode(g[name + '_y']), g[name + '_zf'], rtol=1e-8, atol=1e-9)
        b.set_noise_var(0.01)
        np.testing.assert_allclose(b.decode(g[name + '_y']), g[name + '_mmse'], rtol=1e-8, atol=1e-9)
        with pytest.raises(ValueError):
            b.set_noise_var(-1)
        with pytest.raises(ValueError):
            b.encode(np.zeros(g[name].shape[1] + 1, dtype=complex))
    for name in ('a22', 'a32', 'a12'):
        a = Alamouti(g[name])
        assert a.getNumberOfLayers() == 1 and a.Nt == 2
        np.testing.assert_allclose(a.encode(g[name + '_s']), g[name + '_x'], **TOL)
        np.testing.assert_allclose(a.decode(g[name + '_y']), g[name + '_dec'], **TOL)
        assert abs(a.calc_linear_SINRs(0.1) - np.linalg.norm(g[name], 'fro') ** 2 / 0.1) < 1e-9
    # the reference's encode table (tests/mimo_package_test.py:610-626)
    data = np.r_[0:16] + np.r_[0:16] * 1j
    enc = Alamouti().encode(data)
    assert np.allclose(enc[0, :4] * np.sqrt(2), [0, -1 + 1j, 2 + 2j, -3 + 3j])
    assert np.allclose(enc[1, :4] * np.sqrt(2), [1 + 1j, 0, 3 + 3j, 2 - 2j])
    with pytest.raises(ValueError):
        Alamouti().set_channel_matrix(np.ones((4, 3)))
    with pytest.warns(UserWarning):
        Blast(np.ones((2, 3), dtype=complex))


# ------------------------------------------------------------------ SimulationRunner on the GPU
def test_runner_reference_app_style_and_batched():
    """(i) apps/awgn_modulators/simulate_psk.py's _run_simulation body, verbatim apart from the
    import line, on the façade classes; (ii) the batched form: one fused-link call per repetition."""
    from pyphysim_b200 import links
    from pyphysim_b200.modulators import fundamental
    from pyphysim_b200.simulations import Result, SimulationResults, SimulationRunner, counters_to_results
    from pyphysim_b200.util import misc
    from pyphysim_b200.util.conversion import dB2Linear

    class VerySimplePskSimulationRunner(SimulationRunner):
        def __init__(self):
            super().__init__(read_command_line_args=False)
            self.modulator = fundamental.PSK(4)
            self.NSymbs = 5000
            self.rep_max = 20
            self.max_bit_errors = 1. / 100. * self.NSymbs * self.rep_max
            self.update_progress_function_style = None
            self.params.add('SNR', np.array([0, 6, 12]))
            self.params.set_unpack_parameter('SNR')

        def _run_simulation(self, current_parameters):
            NSymbs = self.NSymbs
            M = self.modulator.M
            SNR = current_parameters["SNR"]
            inputData = np.random.randint(0, M, NSymbs)
            modulatedData = self.modulator.modulate(inputData)
            noiseVar = 1. / dB2Linear(SNR)
            noise = misc.randn_c(NSymbs) * np.sqrt(noiseVar)
            receivedData = modulatedData + noise
            demodulatedData = self.modulator.demodulate(receivedData)
            symbolErrors = sum(inputData != demodulatedData)
            bitErrors = misc.count_bit_errors(inputData, demodulatedData)
            numSymbols = inputData.size
            numBits = inputData.size * fundamental.level2bits(M)
            simResults = SimulationResults()
            simResults.add_result(Result.create("symbol_errors", Result.SUMTYPE, symbolErrors))
            simResults.add_result(Result.create("num_symbols", Result.SUMTYPE, numSymbols))
            simResults.add_result(Result.create("bit_errors", Result.SUMTYPE, bitErrors))
            simResults.add_result(Result.create("num_bits", Result.SUMTYPE, numBits))
            simResults.add_result(Result.create("ber", Result.RATIOTYPE, bitErrors, numBits))
            simResults.add_result(Result.create("ser", Result.RATIOTYPE, symbolErrors, numSymbols))
            return simResults

        def _keep_going(self, current_parameters, simulation_results, current_rep):
            return simulation_results['bit_errors'][-1].get_result() < self.max_bit_errors

    r = VerySimplePskSimulationRunner()
    r.simulate()
    ser = np.array(r.results.get_result_values_list('ser'))
    theory = r.modulator.calcTheoreticalSER(np.array([0., 6., 12.]))
    assert np.all(np.abs(ser - theory) < 0.15 * theory + 2e-4)
    assert r.runned_reps[0] < r.runned_reps[-1] == 20           # early stop at low SNR only

    class Batched(SimulationRunner):
        def __init__(self):
            super().__init__(read_command_line_args=False)
            self.modulator = fundamental.QAM(64)
            self.batch, self.rep_max = 200000, 5
            self.rep = 0
            self.update_progress_function_style = None
            self.params.add('SNR', np.array([10., 20.]))
            self.params.set_unpack_parameter('SNR')

        def _on_simulate_current_params_start(self, current_params):
            self.rep = 0

        def _run_simulation(self, current_parameters):
            c = links.link_siso_flat(self.modulator, 1 / dB2Linear(current_parameters['SNR']), self.batch,
                                     first_unit=self.rep * self.batch, rayleigh=True)
            self.rep += 1
            return counters_to_results(c)

    b = Batched()
    b.simulate()
    assert b.results.get_result_values_list('num_symbols') == [1000000, 1000000]
    ser = b.results.get_result_values_list('ser')
    assert ser[0] > ser[1] > 0.01
    # the same 1e6 realizations in one call give the same counters (batch-size invariance)
    c = links.link_siso_flat(b.modulator, 1 / dB2Linear(20.), 1000000)
    assert c[0] == b.results['symbol_errors'][1].get_result()
