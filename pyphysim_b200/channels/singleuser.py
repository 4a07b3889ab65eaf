"""Single-user channel wrappers with the API of pyphysim.channels.singleuser (thin wrappers around
TdlChannel + a path-loss scale; reference: singleuser.py:19-359)."""
import math

import numpy as np

from . import fading
from .fading_generators import RayleighSampleGenerator

__all__ = ['SuChannel', 'SuMimoChannel']


class SuChannel:
    """SISO channel: TDL channel (default: flat Rayleigh, Ts=1) scaled by sqrt(pathloss)."""

    def __init__(self, fading_generator=None, channel_profile=None, tap_powers_dB=None,
                 tap_delays=None, Ts=None):
        if fading_generator is None:
            fading_generator = RayleighSampleGenerator()
            if channel_profile is None and Ts is None:
                Ts = 1.0
        if channel_profile is None and tap_powers_dB is None and tap_delays is None:
            self._tdlchannel = fading.TdlChannel(fading_generator, tap_powers_dB=np.zeros(1),
                                                 tap_delays=np.zeros(1), Ts=Ts)
        else:
            self._tdlchannel = fading.TdlChannel(fading_generator, channel_profile, tap_powers_dB,
                                                 tap_delays, Ts)
        self._pathloss_value = None

    def set_pathloss(self, pathloss_value=None):
        """singleuser.py:92-108."""
        if pathloss_value is not None:
            if pathloss_value < 0 or pathloss_value > 1:
                raise ValueError("Pathloss must be between 0 and 1")
        self._pathloss_value = pathloss_value

    def set_num_antennas(self, num_rx_antennas, num_tx_antennas):
        self._tdlchannel.set_num_antennas(num_rx_antennas, num_tx_antennas)

    def _scale(self, output):
        if self._pathloss_value is None:
            return output
        k = math.sqrt(self._pathloss_value)
        if isinstance(output, np.ndarray):
            output *= k
            return output
        fading._scale_rows(output, [k])
        return output

    def corrupt_data(self, signal):
        """singleuser.py:130-151."""
        return self._scale(self._tdlchannel.corrupt_data(signal))

    def corrupt_data_in_freq_domain(self, signal, fft_size, carrier_indexes=None):
        return self._scale(self._tdlchannel.corrupt_data_in_freq_domain(signal, fft_size, carrier_indexes))

    def get_last_impulse_response(self):
        """singleuser.py:215-236: impulse response including the path loss."""
        ir = self._tdlchannel.get_last_impulse_response()
        if self._pathloss_value is None:
            return ir
        return math.sqrt(self._pathloss_value) * ir

    @property
    def switched_direction(self):
        return self._tdlchannel.switched_direction

    @switched_direction.setter
    def switched_direction(self, value):
        self._tdlchannel.switched_direction = value

    num_taps = property(lambda self: self._tdlchannel.num_taps)
    num_taps_with_padding = property(lambda self: self._tdlchannel.num_taps_with_padding)
    channel_profile = property(lambda self: self._tdlchannel.channel_profile)
    num_tx_antennas = property(lambda self: self._tdlchannel.num_tx_antennas)
    num_rx_antennas = property(lambda self: self._tdlchannel.num_rx_antennas)


class SuMimoChannel(SuChannel):
    """singleuser.py:305-359."""

    def __init__(self, num_antennas, fading_generator=None, channel_profile=None, tap_powers_dB=None,
                 tap_delays=None, Ts=None):
        if fading_generator is None:
            fading_generator = RayleighSampleGenerator(shape=(num_antennas, num_antennas))
            if channel_profile is None and Ts is None:
                Ts = 1.0
        else:
            fading_generator.shape = (num_antennas, num_antennas)
        super().__init__(fading_generator, channel_profile, tap_powers_dB, tap_delays, Ts)
