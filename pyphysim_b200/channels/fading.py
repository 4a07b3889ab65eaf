"""Tapped-delay-line channel with the API of pyphysim.channels.fading; per-sample work on the GPU.

Host side (one-off): channel profiles and their discretisation.
Device side: fading-sample generation (``b200phy_jakes`` / Philox Rayleigh), the time-varying sparse
FIR (``b200phy_tdl_apply``) and the frequency response (``b200phy_tdl_freq_response``).
"""
import math

import numpy as np

from .. import _device as D
from .. import _lib
from ..util.conversion import dB2Linear, linear2dB
from . import fading_generators as fg

__all__ = ['TdlChannelProfile', 'TdlImpulseResponse', 'TdlChannel', 'TdlMimoChannel',
           'COST259_TUx', 'COST259_RAx', 'COST259_HTx']


class TdlChannelProfile:
    """Powers and delays of the taps of a TDL profile (reference: channels/fading.py:28-316)."""

    def __init__(self, tap_powers_dB=None, tap_delays=None, name='custom'):
        self._name = name
        if tap_powers_dB is None and tap_delays is None:
            tap_powers_dB = np.zeros(1)
            tap_delays = np.zeros(1)
        self._tap_powers_dB = np.array(tap_powers_dB, dtype=float)
        self._tap_powers_dB.flags['WRITEABLE'] = False
        self._tap_powers_linear = dB2Linear(self._tap_powers_dB)
        self._tap_powers_linear.flags['WRITEABLE'] = False
        self._tap_delays = np.array(tap_delays)
        self._tap_delays.flags['WRITEABLE'] = False
        self._num_taps = self._tap_delays.size
        p, d = self._tap_powers_linear, self._tap_delays
        self._mean_excess_delay = float(np.sum(p * d) / np.sum(p))
        self._rms_delay_spread = math.sqrt(float(np.sum(p * d ** 2) / np.sum(p))
                                           - self._mean_excess_delay ** 2)
        self._Ts = None

    mean_excess_delay = property(lambda self: self._mean_excess_delay)
    rms_delay_spread = property(lambda self: self._rms_delay_spread)
    name = property(lambda self: self._name)
    tap_powers_dB = property(lambda self: self._tap_powers_dB)
    tap_powers_linear = property(lambda self: self._tap_powers_linear)
    tap_delays = property(lambda self: self._tap_delays)
    num_taps = property(lambda self: self._num_taps)
    Ts = property(lambda self: self._Ts)

    @property
    def num_taps_with_padding(self):
        if self._Ts is None:
            raise RuntimeError('TdlChannelProfile is not discretized')
        return int(self._tap_delays[-1] + 1)

    @property
    def is_discretized(self):
        return self._Ts is not None

    def get_discretize_profile(self, Ts):
        """fading.py:236-304: round delays to samples, sum colliding taps (linear), normalise."""
        if self.is_discretized:
            raise RuntimeError("Trying to discretize a TdlChannelProfile "
                               "object that is already discretized.")
        powers, delays = self._calc_discretized_tap_powers_and_delays(Ts)
        out = TdlChannelProfile(powers, delays, "{0} (discretized)".format(self.name))
        out._Ts = Ts
        return out

    def _calc_discretized_tap_powers_and_delays(self, Ts):
        slots = np.round(self._tap_delays / Ts).astype(int).flatten()
        delay_indexes, inverse = np.unique(slots, return_inverse=True)
        acc = np.zeros(delay_indexes.size)
        np.add.at(acc, inverse, self._tap_powers_linear)
        acc /= np.sum(acc)
        return linear2dB(acc), delay_indexes

    def __repr__(self):
        return "<TdlChannelProfile: '{0}' ({1} taps)>".format(self.name, self.num_taps)


# 3GPP TR 25.943 (COST 259) profiles, values as tabulated in channels/fading.py:327-353
COST259_TUx = TdlChannelProfile(
    np.array([-5.7, -7.6, -10.1, -10.2, -10.2, -11.5, -13.4, -16.3, -16.9, -17.1,
              -17.4, -19, -19, -19.8, -21.5, -21.6, -22.1, -22.6, -23.5, -24.3]),
    np.array([0, 217, 512, 514, 517, 674, 882, 1230, 1287, 1311, 1349, 1533, 1535,
              1622, 1818, 1836, 1884, 1943, 2048, 2140]) * 1e-9, 'COST259_TU')
COST259_RAx = TdlChannelProfile(
    np.array([-5.2, -6.4, -8.4, -9.3, -10.0, -13.1, -15.3, -18.5, -20.4, -22.4]),
    np.array([0., 42., 101., 129., 149., 245., 312., 410., 469., 528]) * 1e-9, 'COST259_RA')
COST259_HTx = TdlChannelProfile(
    np.array([-3.6, -8.9, -10.2, -11.5, -11.8, -12.7, -13.0, -16.2, -17.3, -17.7,
              -17.6, -22.7, -24.1, -25.8, -25.8, -26.2, -29.0, -29.9, -30.0, -30.7]),
    np.array([0., 356., 441., 528., 546., 609., 625., 842., 916., 941., 15000.,
              16172., 16492., 16876., 16882., 16978., 17615., 17827., 17849., 18016.]) * 1e-9,
    'COST259_HT')
