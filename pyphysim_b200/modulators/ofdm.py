"""OFDM modulator / one-tap equaliser with the API of pyphysim.modulators.ofdm.

Parameter checks and subcarrier maps are host integers; the IFFT/FFT with
scatter/gather and cyclic prefix, and the equaliser, run on the GPU (``b200phy_ofdm_mod``,
``b200phy_ofdm_demod``, ``b200phy_ofdm_equalize``).
"""
import ctypes as C

import numpy as np

from .. import _device as D
from .. import _lib

__all__ = ['OFDM', 'OfdmOneTapEqualizer']


class OFDM:
    """reference: modulators/ofdm.py:20-466."""

    def __init__(self, fft_size, cp_size, num_used_subcarriers=None):
        self.fft_size = 0
        self.cp_size = 0
        self.num_used_subcarriers = 0
        self.set_parameters(fft_size, cp_size, num_used_subcarriers)

    def set_parameters(self, fft_size, cp_size, num_used_subcarriers=None):
        """ofdm.py:56-94."""
        if (cp_size < 0) or cp_size > fft_size:
            raise ValueError("cp_size must be nonnegative and cannot be greater than fft_size")
        if num_used_subcarriers is None:
            num_used_subcarriers = fft_size
        if num_used_subcarriers > fft_size:
            raise ValueError("Number of used subcarriers cannot be greater than the fft_size")
        if (num_used_subcarriers % 2 != 0) or (num_used_subcarriers < 2):
            raise ValueError("Number of used subcarriers must be a multiple of 2")
        self.fft_size = fft_size
        self.cp_size = cp_size
        self.num_used_subcarriers = num_used_subcarriers

    def _calc_zeropad(self, input_data_size):
        """ofdm.py:96-123."""
        num_ofdm_symbols = int(np.ceil(float(input_data_size) / self.num_used_subcarriers))
        return self.num_used_subcarriers * num_ofdm_symbols - input_data_size, num_ofdm_symbols

    def _get_subcarrier_numbers(self):
        """ofdm.py:125-151: 0..fft/2-1 followed by -fft/2..-1."""
        n = np.arange(self.fft_size)
        return np.where(n < self.fft_size // 2 + self.fft_size % 2, n, n - self.fft_size)

    def _get_used_subcarrier_numbers(self):
        """ofdm.py:153-186."""
        if self.num_used_subcarriers == self.fft_size:
            return self._get_subcarrier_numbers()
        half = self.num_used_subcarriers // 2
        return np.concatenate([np.arange(1, half + 1), np.arange(-half, 0)])

    def get_used_subcarrier_indexes(self):
        """ofdm.py:188-224: FFT bin of every data position."""
        numbers = self._get_used_subcarrier_numbers()
        half = self.num_used_subcarriers // 2
        return np.concatenate([self.fft_size + numbers[half:], numbers[0:half]])

    def _calculate_power_scale(self):
        """ofdm.py:370-392."""
        return (float(self.fft_size) ** 2) / (float(self.num_used_subcarriers) + self.cp_size)

    def _prepare_input_signal(self, input_signal):
        """ofdm.py:226-281 (host helper kept for API compatibility; modulate() does this on the GPU)."""
        input_signal = np.asarray(input_signal)
        zeropad, n_sym = self._calc_zeropad(input_signal.size)
        grid = np.zeros([n_sym, self.fft_size], dtype=complex)
        padded = np.concatenate([input_signal.reshape(-1), np.zeros(zeropad)])
        grid[:, self.get_used_subcarrier_indexes()] = padded.reshape(n_sym, self.num_used_subcarriers)
        return grid

    def _dtype(self, x):
        return D.dtype_of_samples(x) if hasattr(x, 'dtype') else _lib.F64

    def modulate(self, input_signal):
        """ofdm.py:394-429: zero-pad, scatter, scaled IFFT, cyclic prefix, flatten."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        dtype = self._dtype(input_signal)
        if dtype == _lib.F64 and not D.is_torch(input_signal):
            input_signal = np.asarray(input_signal).astype(complex, copy=False)
        x, was_np = D.to_device(input_signal, D.complex_np(dtype))
        x = x.reshape(-1)
        zeropad, n_sym = self._calc_zeropad(x.numel())
        if zeropad:
            x = torch.cat([x, torch.zeros(zeropad, dtype=x.dtype, device=x.device)])
        out = torch.empty(n_sym * (self.fft_size + self.cp_size), dtype=x.dtype, device=x.device)
        _lib.check(lib.b200phy_ofdm_mod(dtype, _lib.ptr(x), _lib.ptr(out), 1, n_sym, self.fft_size,
                                        self.cp_size, self.num_used_subcarriers, _lib.cur_stream()))
        return D.from_device(out, was_np)

    def demodulate(self, received_signal):
        """ofdm.py:431-466: strip CP, FFT / sqrt(scale), gather used bins (zero-pad is not removed).
        Like the reference this reshapes a NumPy input in place (ofdm.py:363-366)."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        dtype = self._dtype(received_signal)
        r, was_np = D.to_device(received_signal, D.complex_np(dtype))
        n_sym = r.numel() // (self.fft_size + self.cp_size)
        if isinstance(received_signal, np.ndarray):
            received_signal.shape = (n_sym, self.fft_size + self.cp_size)
        r = r.reshape(-1)
        out = torch.empty(n_sym * self.num_used_subcarriers, dtype=r.dtype, device=r.device)
        _lib.check(lib.b200phy_ofdm_demod(dtype, _lib.ptr(r), _lib.ptr(out), 1, n_sym, self.fft_size,
                                          self.cp_size, self.num_used_subcarriers, _lib.cur_stream()))
        return D.from_device(out, was_np)


class OfdmOneTapEqualizer:
    """reference: modulators/ofdm.py:469-552."""

    def __init__(self, ofdm_obj):
        self._ofdm_obj = ofdm_obj

    def equalize_data(self, data, impulse_response):
        """ofdm.py:515-552: y / H with H the frequency response averaged over the samples (CP
        included) of each OFDM symbol — evaluated as DFT(mean taps), which is the same by linearity."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        o = self._ofdm_obj
        y, was_np = D.to_device(data, np.complex128)
        y = y.reshape(-1)
        n_sym = y.numel() // o.num_used_subcarriers
        taps = impulse_response._dev()
        if taps.dim() != 2:
            raise ValueError("OfdmOneTapEqualizer needs a SISO impulse response")
        if impulse_response.num_samples != n_sym * (o.fft_size + o.cp_size):
            raise ValueError("impulse response has %d samples, expected %d"
                             % (impulse_response.num_samples, n_sym * (o.fft_size + o.cp_size)))
        delays = np.ascontiguousarray(impulse_response.tap_indexes_sparse, dtype=np.int32)
        ones = np.ones(delays.size)
        out = torch.empty_like(y)
        _lib.check(lib.b200phy_ofdm_equalize(_lib.F64, _lib.ptr(y), _lib.ptr(taps),
                                             ones.ctypes.data_as(C.POINTER(C.c_double)),
                                             delays.ctypes.data_as(C.POINTER(C.c_int32)), delays.size,
                                             1, 1, n_sym, o.fft_size, o.cp_size, o.num_used_subcarriers, 0.0,
                                             _lib.ptr(out), _lib.cur_stream()))
        return D.from_device(out, was_np)
