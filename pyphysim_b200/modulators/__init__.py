"""Mirror of pyphysim.modulators: digital modulators and OFDM."""
from .fundamental import BPSK, PSK, QAM, QPSK, Modulator  # noqa: F401
from .ofdm import OFDM, OfdmOneTapEqualizer  # noqa: F401

__all__ = ['Modulator', 'PSK', 'QPSK', 'BPSK', 'QAM', 'OFDM', 'OfdmOneTapEqualizer']
