"""The `pyphysim` drop-in namespace (repo-root package `pyphysim/`): the import lines of the reference's hot-path
simulators (apps/awgn_modulators/simulate_psk.py:1-19, apps/mimo/simulate_mimo.py:1-20,
apps/ofdm/ofdm_tdlchannel.py:1-12) resolve unchanged, to the SAME objects as pyphysim_b200, and expose every
name SURVEY.md §8(b) lists."""
import importlib

import pytest

SURFACE = {
    'pyphysim.modulators': ['Modulator', 'PSK', 'QPSK', 'BPSK', 'QAM', 'OFDM'],
    'pyphysim.modulators.fundamental': ['Modulator', 'PSK', 'QPSK', 'BPSK', 'QAM'],
    'pyphysim.modulators.ofdm': ['OFDM', 'OfdmOneTapEqualizer'],
    'pyphysim.channels.fading_generators': ['JakesSampleGenerator', 'RayleighSampleGenerator'],
    'pyphysim.channels.fading': ['TdlChannelProfile', 'COST259_TUx', 'COST259_RAx', 'COST259_HTx',
                                 'TdlImpulseResponse', 'TdlChannel', 'TdlMimoChannel'],
    'pyphysim.channels.singleuser': ['SuChannel', 'SuMimoChannel'],
    'pyphysim.channels.multiuser': ['MuChannel', 'MuMimoChannel'],
    'pyphysim.mimo': ['MimoBase', 'Blast', 'Alamouti', 'MRT', 'MRC', 'SVDMimo', 'GMDMimo'],
    'pyphysim.mimo.mimo': ['MimoBase', 'Blast', 'Alamouti', 'MRT', 'MRC', 'SVDMimo', 'GMDMimo',
                           'calc_post_processing_SINRs'],
    'pyphysim.util.misc': ['randn_c', 'count_bits', 'count_bit_errors', 'level2bits', 'qfunc', 'pretty_time', 'gmd'],
    'pyphysim.util.conversion': ['dB2Linear', 'linear2dB', 'binary2gray', 'gray2binary', 'dBm2Linear',
                                 'linear2dBm', 'SNR_dB_to_EbN0_dB', 'EbN0_dB_to_SNR_dB'],
    'pyphysim.simulations': ['SimulationRunner', 'SimulationParameters', 'SimulationResults', 'Result', 'SkipThisOne'],
    'pyphysim.simulations.runner': ['SimulationRunner', 'SkipThisOne', 'get_partial_results_filename'],
    'pyphysim.simulations.results': ['Result', 'SimulationResults'],
    'pyphysim.simulations.parameters': ['SimulationParameters'],
    'pyphysim.reference_signals.zadoffchu': ['calcBaseZC', 'get_extended_ZF'],
    'pyphysim.channel_estimation.estimators': ['compute_ls_estimation', 'compute_mmse_estimation'],
}


@pytest.mark.parametrize('module', sorted(SURFACE))
def test_module_resolves_and_exposes_the_reference_names(module):
    mod = importlib.import_module(module)
    real = importlib.import_module('pyphysim_b200' + module[len('pyphysim'):])
    assert mod is real                                    # an alias, not a second copy
    missing = [n for n in SURFACE[module] if not hasattr(mod, n)]
    assert not missing, '%s lacks %s' % (module, missing)


def test_reference_app_import_lines_run_unchanged():
    ns = {}
    exec('\n'.join([
        'from pyphysim.modulators import fundamental',
        'from pyphysim.modulators import OFDM, QPSK',
        'from pyphysim.modulators.ofdm import OFDM',
        'from pyphysim.modulators.ofdm import OfdmOneTapEqualizer',
        'from pyphysim.channels.singleuser import SuChannel',
        'from pyphysim.channels.fading_generators import JakesSampleGenerator',
        'from pyphysim.channels.fading import COST259_TUx',
        'from pyphysim.util.conversion import dB2Linear, linear2dB',
        'from pyphysim.util import misc',
        'from pyphysim.simulations import (Result, SimulationParameters, SimulationResults, SimulationRunner)',
        'from pyphysim.simulations import *',
        'from pyphysim.mimo import mimo',
    ]), ns)
    assert ns['fundamental'].QAM(16).M == 16 and ns['dB2Linear'](10.0) == 10.0
    assert issubclass(ns['mimo'].Alamouti, ns['mimo'].MimoBase)
    assert 'SimulationRunner' in ns and 'SkipThisOne' in ns


def test_out_of_scope_packages_say_so():
    import pyphysim
    with pytest.raises(ModuleNotFoundError):
        importlib.import_module('pyphysim.ia')
    with pytest.raises(ModuleNotFoundError, match='OUT OF SCOPE'):
        pyphysim.comm
