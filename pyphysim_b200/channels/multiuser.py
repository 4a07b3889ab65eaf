"""Multi-link channel grids with the API of pyphysim.channels.multiuser: ``MuChannel`` (a num_rx x num_tx
grid of independent SISO ``SuChannel`` links) and ``MuMimoChannel`` (the same grid with MIMO links) —
SURVEY.md §8f row next-2, reference multiuser.py:42-586.

The grid is orchestration only: every link is a ``SuChannel`` whose fading generation, TDL convolution
(``b200phy_jakes`` / ``b200phy_tdl_apply``) and block-static frequency-domain application
(``b200phy_tdl_freq_response`` + ``b200phy_freq_apply``) run on the GPU; the receivers' sums are device
adds when the inputs are device tensors.  ``MultiUserChannelMatrix`` (block channel with SINR / covariance
algebra, multiuser.py:586-2807) belongs to the interference-alignment workloads and is out of scope.
"""
import numpy as np

from . import singleuser
from .fading_generators import RayleighSampleGenerator

__all__ = ['MuChannel', 'MuMimoChannel']


class MuChannel:
    """num_rx x num_tx grid of SISO links, all with the same power delay profile (multiuser.py:42-395)."""

    def __init__(self, N, fading_generator=None, channel_profile=None, tap_powers_dB=None, tap_delays=None,
                 Ts=None):
        if fading_generator is None:
            fading_generator = RayleighSampleGenerator()
        num_rx, num_tx = N if isinstance(N, (tuple, list)) else (N, N)
        self._su_siso_channels = np.empty((num_rx, num_tx), dtype=object)
        for rx in range(num_rx):
            for tx in range(num_tx):
                # every link fades independently; the first link's (possibly discretised) profile is
                # reused by all others (multiuser.py:121-141)
                link = singleuser.SuChannel(fading_generator.get_similar_fading_generator(),
                                            channel_profile=channel_profile, tap_powers_dB=tap_powers_dB,
                                            tap_delays=tap_delays, Ts=Ts)
                self._su_siso_channels[rx, tx] = link
                channel_profile = link.channel_profile
        self._pathloss_matrix = None

    def __repr__(self):
        return "{0}(shape={1}x{2}, switched={3})".format(self.__class__.__name__, *self._su_siso_channels.shape,
                                                         self.switched_direction)

    @property
    def switched_direction(self):
        return self._su_siso_channels[0, 0].switched_direction

    @switched_direction.setter
    def switched_direction(self, value):
        for link in self._su_siso_channels.flat:
            link.switched_direction = value

    @property
    def num_tx_antennas(self):
        """Antennas of every transmitter (multiuser.py:170-187)."""
        return np.array([link.num_tx_antennas for link in self._su_siso_channels[0, :]], dtype=int)

    @property
    def num_rx_antennas(self):
        return np.array([link.num_rx_antennas for link in self._su_siso_channels[:, 0]], dtype=int)

    channel_profile = property(lambda self: self._su_siso_channels[0, 0].channel_profile)
    num_taps = property(lambda self: self._su_siso_channels[0, 0].num_taps)
    num_taps_with_padding = property(lambda self: self._su_siso_channels[0, 0].num_taps_with_padding)
    pathloss_matrix = property(lambda self: self._pathloss_matrix)

    def set_pathloss(self, pathloss_matrix=None):
        """multiuser.py:231-252 (None removes the path loss)."""
        num_rx, num_tx = self._su_siso_channels.shape
        self._pathloss_matrix = None if pathloss_matrix is None else np.copy(pathloss_matrix)
        for rx in range(num_rx):
            for tx in range(num_tx):
                self._su_siso_channels[rx, tx].set_pathloss(None if pathloss_matrix is None
                                                            else pathloss_matrix[rx, tx])

    def _links(self):
        return self._su_siso_channels.T if self.switched_direction else self._su_siso_channels

    def _through(self, signal, send):
        links = self._links()
        num_rx, num_tx = links.shape
        if num_tx == 1 and getattr(signal, 'ndim', 0) == 1:
            signal = signal.reshape(1, -1)
        outputs = np.empty(num_rx, dtype=object)
        for rx in range(num_rx):
            acc = send(links[rx, 0], signal[0])
            for tx in range(1, num_tx):
                acc = acc + send(links[rx, tx], signal[tx])
            outputs[rx] = acc
        return outputs

    def corrupt_data(self, signal):
        """multiuser.py:254-300: signal[tx] through link (rx, tx), summed per receiver; returns an object
        array with one entry per receiver."""
        return self._through(signal, lambda link, s: link.corrupt_data(s))

    def corrupt_data_in_freq_domain(self, signal, fft_size, carrier_indexes=None):
        """multiuser.py:302-360: the block-static frequency-domain variant."""
        return self._through(signal, lambda link, s: link.corrupt_data_in_freq_domain(s, fft_size, carrier_indexes))

    def get_last_impulse_response(self, rx_idx, tx_idx):
        """multiuser.py:362-395."""
        return self._su_siso_channels[rx_idx, tx_idx].get_last_impulse_response()


class MuMimoChannel(MuChannel):
    """The grid with num_rx_antennas x num_tx_antennas MIMO links (multiuser.py:398-583)."""

    def __init__(self, N, num_rx_antennas, num_tx_antennas, fading_generator=None, channel_profile=None,
                 tap_powers_dB=None, tap_delays=None, Ts=None):
        super().__init__(N, fading_generator, channel_profile, tap_powers_dB, tap_delays, Ts)
        for link in self._su_siso_channels.flat:
            link.set_num_antennas(num_rx_antennas, num_tx_antennas)
