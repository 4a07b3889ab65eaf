"""`pyphysim` drop-in namespace: the reference's module paths, served by pyphysim_b200.

The simulators of the reference import `pyphysim.*` (apps/awgn_modulators/simulate_psk.py:1-19,
apps/mimo/simulate_mimo.py:1-20, apps/ofdm/ofdm_tdlchannel.py:1-12, the notebooks).  With this directory on
sys.path (it sits at the repo root next to pyphysim_b200/) those import lines resolve unchanged:

    from pyphysim.modulators import fundamental          from pyphysim.util import misc
    from pyphysim.modulators.ofdm import OFDM, OfdmOneTapEqualizer
    from pyphysim.channels.fading import COST259_TUx    from pyphysim.mimo import mimo
    from pyphysim.simulations import *                   from pyphysim.util.conversion import dB2Linear

Every `pyphysim.X` is the SAME module object as `pyphysim_b200.X` (an alias in sys.modules, not a second copy):
classes keep one identity, the CUDA library is loaded once, and result files pickled by the reference
(class paths `pyphysim.simulations.results.SimulationResults`, ...) unpickle into this implementation.
Only the hot-path packages of SURVEY.md §8 exist; `import pyphysim.ia` etc. raise ModuleNotFoundError (and the
attribute `pyphysim.ia` says why).
"""
import importlib
import sys

import pyphysim_b200 as _impl

__version__ = _impl.__version__

_SUBMODULES = (
    'util', 'util.misc', 'util.conversion',
    'modulators', 'modulators.fundamental', 'modulators.ofdm',
    'channels', 'channels.fading_generators', 'channels.fading', 'channels.singleuser', 'channels.multiuser',
    'mimo', 'mimo.mimo',
    'simulations', 'simulations.parameters', 'simulations.results', 'simulations.runner',
    'reference_signals', 'reference_signals.zadoffchu', 'reference_signals.root_sequence',
    'reference_signals.srs', 'reference_signals.dmrs', 'reference_signals.channel_estimation',
    'channel_estimation', 'channel_estimation.estimators',
)

for _name in _SUBMODULES:
    _mod = importlib.import_module('pyphysim_b200.' + _name)
    sys.modules[__name__ + '.' + _name] = _mod
    if '.' not in _name:
        globals()[_name] = _mod

_OUT_OF_SCOPE = ('ia', 'comm', 'cell', 'subspace', 'pointprocess', 'extra', 'progressbar', 'c_extensions')


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise ModuleNotFoundError("pyphysim.%s is outside the hot-path scope of pyphysim_b200 (SURVEY.md §2 "
                                  "OUT OF SCOPE); use the reference package for it" % name)
    raise AttributeError("module 'pyphysim' has no attribute %r" % name)
