"""MIMO schemes of the hot path with the API of pyphysim.mimo.mimo: Blast (ZF / MMSE) and Alamouti.

Encode / decode run on the GPU (``b200phy_blast_decode``, ``b200phy_alamouti_encode/decode``);
the channel matrix itself is a tiny host array (Nr x Nt <= 4 x 4).  MRT / MRC / SVDMimo / GMDMimo and
the post-processing SINR helpers are SURVEY.md §8f row next-3 and not built yet.
"""
import math
import warnings

import numpy as np

from .. import _device as D
from .. import _lib

__all__ = ['MimoBase', 'Blast', 'Alamouti']


class MimoBase:
    """reference: mimo/mimo.py:117-462 (the parts Blast and Alamouti need)."""

    def __init__(self, channel=None):
        self._channel = None
        if channel is not None:
            self.set_channel_matrix(channel)

    def set_channel_matrix(self, channel):
        """mimo.py:171-190 (a 1-D array is a single-receive-antenna channel)."""
        channel = np.asarray(channel)
        if channel.ndim == 1:
            channel = channel[np.newaxis, :]
        self._channel = channel.astype(complex)

    @property
    def Nt(self):
        return self._channel.shape[1]

    @property
    def Nr(self):
        return self._channel.shape[0]

    def getNumberOfLayers(self):  # pragma: no cover
        raise NotImplementedError("getNumberOfLayers still needs to be implemented in the subclass")

    def encode(self, transmit_data):  # pragma: no cover
        raise NotImplementedError("encode still needs to be implemented in the subclass")

    def decode(self, received_data):  # pragma: no cover
        raise NotImplementedError("decode still needs to be implemented in the subclass")

    def _channel_dev(self):
        t, _ = D.to_device(self._channel, np.complex128)
        return t


class Blast(MimoBase):
    """Spatial multiplexing with a linear ZF / MMSE receiver (mimo.py:465-660)."""

    def __init__(self, channel=None):
        super().__init__(channel)
        self._noise_var = 0.0

    def set_channel_matrix(self, channel):
        Nr, Nt = np.asarray(channel).shape
        if Nt > Nr:
            warnings.warn("The number of transmit antennas for {0} should not be greater than the "
                          "number of receive antennas.".format(self.__class__.__name__))
        super().set_channel_matrix(channel)

    def getNumberOfLayers(self):
        return self.Nt

    def set_noise_var(self, noise_var):
        """mimo.py:531-553: None or 0 -> zero forcing, > 0 -> MMSE."""
        if noise_var is None:
            self._noise_var = 0.0
        elif noise_var >= 0.0:
            self._noise_var = noise_var
        else:
            raise ValueError('Noise variance must be a non-negative value.')

    def encode(self, transmit_data):
        """mimo.py:609-640: symbol k goes to antenna k mod Nt, power split 1/sqrt(Nt)."""
        from ..channels.fading import _scale_rows
        x, was_np = D.to_device(transmit_data, np.complex128)
        n_streams = self.getNumberOfLayers()
        if x.numel() % n_streams != 0:
            raise ValueError("Input array number of elements must be a multiple of the"
                             " number of transmit antennas")
        enc = x.reshape(-1, n_streams).t().contiguous()            # reshape(nStreams, -1, order='F')
        _scale_rows(enc, [1.0 / math.sqrt(self.Nt)])
        return D.from_device(enc, was_np)

    def decode(self, received_data):
        """mimo.py:642-660."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        y, was_np = D.to_device(received_data, np.complex128)
        if y.dim() == 1:
            y = y.reshape(-1, 1)
        Nr, T = y.shape
        if Nr != self.Nr:
            raise ValueError("received_data must have %d rows" % self.Nr)
        H = self._channel_dev()
        out = torch.empty(self.Nt * T, dtype=torch.complex128, device='cuda')
        _lib.check(lib.b200phy_blast_decode(_lib.F64, _lib.ptr(H), _lib.ptr(y.contiguous()), 1, self.Nr,
                                            self.Nt, T, float(self._noise_var), _lib.ptr(out),
                                            _lib.cur_stream()))
        return D.from_device(out, was_np)


class Alamouti(MimoBase):
    """Alamouti space-time block code, 2 transmit antennas (mimo.py:1073-1287)."""

    def set_channel_matrix(self, channel):
        channel = np.asarray(channel)
        if channel.ndim == 1:
            super().set_channel_matrix(channel[np.newaxis, :])
        else:
            _, Nt = channel.shape
            if Nt != 2:
                raise ValueError("The number of transmit antennas must be equal to 2 for the "
                                 "{0} scheme".format(self.__class__.__name__))
            super().set_channel_matrix(channel)

    def getNumberOfLayers(self):
        return 1

    def calc_linear_SINRs(self, noise_var):
        """mimo.py:1133-1164: ||H||_F^2 / noise_var (host scalar)."""
        return np.linalg.norm(self._channel, 'fro') ** 2 / noise_var

    def calc_SINRs(self, noise_var):
        return 10.0 * np.log10(self.calc_linear_SINRs(noise_var))

    def encode(self, transmit_data):
        """mimo.py:1166-1214."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        s, was_np = D.to_device(transmit_data, np.complex128)
        s = s.reshape(-1)
        T = s.numel()
        x = torch.empty((2, T), dtype=torch.complex128, device='cuda')
        _lib.check(lib.b200phy_alamouti_encode(_lib.F64, _lib.ptr(s), 1, T, _lib.ptr(x), _lib.cur_stream()))
        return D.from_device(x, was_np)

    def decode(self, received_data):
        """mimo.py:1216-1287."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        y, was_np = D.to_device(received_data, np.complex128)
        if y.dim() == 1:
            y = y.reshape(1, -1)
        Nr, T = y.shape
        H = self._channel_dev()
        out = torch.empty(T, dtype=torch.complex128, device='cuda')
        _lib.check(lib.b200phy_alamouti_decode(_lib.F64, _lib.ptr(H), _lib.ptr(y.contiguous()), 1, Nr, T,
                                               _lib.ptr(out), _lib.cur_stream()))
        return D.from_device(out, was_np)
