"""OFDM modulate / demodulate / one-tap equaliser — NumPy restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference:
``pyphysim/modulators/ofdm.py``.
"""
import math

import numpy as np


def check_parameters(fft_size, cp_size, num_used=None):
    """OFDM.set_parameters (ofdm.py:56-94)."""
    if (cp_size < 0) or cp_size > fft_size:
        raise ValueError("cp_size must be nonnegative and cannot be greater than fft_size")
    if num_used is None:
        num_used = fft_size
    if num_used > fft_size:
        raise ValueError("Number of used subcarriers cannot be greater than the fft_size")
    if (num_used % 2 != 0) or (num_used < 2):
        raise ValueError("Number of used subcarriers must be a multiple of 2")
    return fft_size, cp_size, num_used


def used_subcarrier_indexes(fft_size, num_used):
    """OFDM.get_used_subcarrier_indexes (ofdm.py:125-224): FFT-bin index of
    data position q.  Full allocation uses fftshift numbering (DC used)."""
    if num_used == fft_size:
        numbers = np.fft.fftshift(np.arange(fft_size) - fft_size // 2)
    else:
        half = num_used // 2
        numbers = np.hstack([np.r_[1:half + 1], np.r_[-half:0]])
    half = num_used // 2
    return np.hstack([fft_size + numbers[half:], numbers[0:half]])


def power_scale(fft_size, cp_size, num_used):
    """OFDM._calculate_power_scale (ofdm.py:370-392)."""
    return (float(fft_size) ** 2) / (float(num_used) + cp_size)


def calc_zeropad(n, num_used):
    """OFDM._calc_zeropad (ofdm.py:96-123)."""
    n_sym = int(np.ceil(float(n) / num_used))
    return num_used * n_sym - n, n_sym


def modulate(x, fft_size, cp_size, num_used):
    """OFDM.modulate (ofdm.py:394-429; helpers :226-281, :320-341)."""
    zp, n_sym = calc_zeropad(x.size, num_used)
    x = np.hstack([x, np.zeros(zp)]).reshape(n_sym, num_used)
    grid = np.zeros([n_sym, fft_size], dtype=complex)
    grid[:, used_subcarrier_indexes(fft_size, num_used)] = x
    t = math.sqrt(power_scale(fft_size, cp_size, num_used)) * np.fft.ifft(grid, fft_size, 1)
    if cp_size != 0:
        t = np.hstack([t[:, -cp_size:], t])
    return t.flatten()


def demodulate(r, fft_size, cp_size, num_used):
    """OFDM.demodulate (ofdm.py:431-466; _remove_CP :343-368; zero-pad is NOT
    removed :283-318).  Unlike the reference this does not reshape the caller's
    array in place."""
    n_sym = r.size // (fft_size + cp_size)
    r = np.reshape(r, (n_sym, fft_size + cp_size))[:, cp_size:]
    f = np.fft.fft(r, fft_size, 1) / math.sqrt(power_scale(fft_size, cp_size, num_used))
    return f[:, used_subcarrier_indexes(fft_size, num_used)].flatten()


def dense_taps(tap_values_sparse, delays):
    """TdlImpulseResponse._get_samples_including_the_extra_zeros (fading.py:482-511)."""
    shape = (int(delays[-1]) + 1,) + tap_values_sparse.shape[1:]
    dense = np.zeros(shape, dtype=complex)
    dense[delays] = tap_values_sparse
    return dense


def mean_freq_response_reference(tap_values_sparse, delays, fft_size, n_sym):
    """OfdmOneTapEqualizer.equalize_data (ofdm.py:539-548) exactly as written:
    FFT of the dense taps for EVERY time sample (fading.py:513-536), reshape
    (fft, n_sym, fft+cp), mean over the samples of each OFDM symbol (CP included).
    SISO taps only ([taps, N]).  Returns [n_sym, fft]."""
    H = np.fft.fft(dense_taps(tap_values_sparse, delays), fft_size, axis=0)
    H = np.reshape(H, (fft_size, n_sym, -1))
    return np.mean(H, axis=2).T


def mean_freq_response(tap_values_sparse, delays, fft_size, n_sym):
    """Restatement used by the device kernels: by linearity of the FFT,
    mean_n FFT(h_n) == FFT(mean_n h_n).  Works for SISO [taps, N] and MIMO
    [taps, Nr, Nt, N]; returns [n_sym, fft(, Nr, Nt)].  Pinned against
    ``mean_freq_response_reference`` in tests/test_oracle_golden.py."""
    tv = tap_values_sparse
    N = tv.shape[-1]
    m = tv.reshape(tv.shape[:-1] + (n_sym, N // n_sym)).mean(axis=-1)   # [taps,(Nr,Nt,)n_sym]
    m = np.moveaxis(m, -1, 0)                                           # [n_sym,taps,(Nr,Nt)]
    dense = np.zeros((n_sym, int(delays[-1]) + 1) + m.shape[2:], dtype=complex)
    dense[:, delays] = m
    return np.fft.fft(dense, fft_size, axis=1)


def onetap_equalize(y, Hmean, fft_size, num_used):
    """OfdmOneTapEqualizer._equalize_data (ofdm.py:483-513): y / H[:, used bins]."""
    y = np.reshape(y, (-1, num_used))
    return (y / Hmean[:, used_subcarrier_indexes(fft_size, num_used)]).flatten()
