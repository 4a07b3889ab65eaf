"""Jakes / Rayleigh fading and the tapped-delay-line channel — NumPy restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference:
``pyphysim/channels/fading_generators.py``, ``pyphysim/channels/fading.py``.
"""
import math

import numpy as np

from .modulators import dB2Linear, linear2dB

# 3GPP TR 25.943 COST-259 profiles as tabulated in channels/fading.py:327-353
# (powers in dB, delays in seconds).
COST259_TU = (np.array([-5.7, -7.6, -10.1, -10.2, -10.2, -11.5, -13.4, -16.3, -16.9, -17.1,
                        -17.4, -19, -19, -19.8, -21.5, -21.6, -22.1, -22.6, -23.5, -24.3]),
              np.array([0, 217, 512, 514, 517, 674, 882, 1230, 1287, 1311, 1349, 1533, 1535,
                        1622, 1818, 1836, 1884, 1943, 2048, 2140]) * 1e-9)
COST259_RA = (np.array([-5.2, -6.4, -8.4, -9.3, -10.0, -13.1, -15.3, -18.5, -20.4, -22.4]),
              np.array([0., 42., 101., 129., 149., 245., 312., 410., 469., 528]) * 1e-9)
COST259_HT = (np.array([-3.6, -8.9, -10.2, -11.5, -11.8, -12.7, -13.0, -16.2, -17.3, -17.7,
                        -17.6, -22.7, -24.1, -25.8, -25.8, -26.2, -29.0, -29.9, -30.0, -30.7]),
              np.array([0., 356., 441., 528., 546., 609., 625., 842., 916., 941., 15000.,
                        16172., 16492., 16876., 16882., 16978., 17615., 17827., 17849.,
                        18016.]) * 1e-9)


def discretize_profile(powers_dB, delays, Ts):
    """TdlChannelProfile._calc_discretized_tap_powers_and_delays + the
    constructor round trip (fading.py:272-304, :80-86): colliding taps are summed
    in linear scale, normalised to sum 1, converted to dB and back to linear.
    Returns (tap_powers_linear, tap_delay_indexes)."""
    idx, inv = np.unique(np.round(delays / Ts).astype(int).flatten(), return_inverse=True)
    p = np.zeros(idx.size)
    for i, v in enumerate(dB2Linear(powers_dB)):
        p[inv[i]] += v
    p /= np.sum(p)
    return dB2Linear(linear2dB(p)), idx


def jakes_time(t0, Ts, N):
    """JakesSampleGenerator._generate_time_samples (fading_generators.py:459-467).
    Returns (t[N], new_t0)."""
    t = np.arange(t0, N * Ts + t0, Ts * 1.0000000001)
    return t, t[-1] + Ts


def jakes_samples(phi, psi, Fd, Ts, t0, N):
    """JakesSampleGenerator.generate_more_samples (fading_generators.py:495-523).
    phi/psi have shape (L, *shape); returns (h[*shape, N], new_t0)."""
    L = phi.shape[0]
    t, t1 = jakes_time(t0, Ts, N)
    t = t.reshape([1] * phi.ndim + [N])
    h = math.sqrt(1.0 / L) * np.sum(
        np.exp(1j * (2 * np.pi * Fd * np.cos(phi[..., None]) * t + psi[..., None])), axis=0)
    return h, t1


def tdl_taps(fading_samples, tap_powers_linear):
    """TdlChannel.generate_impulse_response (fading.py:908-959): samples *
    sqrt(P_tap) broadcast over the leading (tap) axis."""
    shape = [tap_powers_linear.size] + [1] * (fading_samples.ndim - 1)
    return fading_samples * np.sqrt(np.reshape(tap_powers_linear, shape))


def tdl_corrupt(signal, taps, delays):
    """TdlChannel.corrupt_data (fading.py:1046-1124), forward direction.
    SISO: signal[N], taps[taps, N]  -> out[N + mem].
    MIMO: signal[Nt, N], taps[taps, Nr, Nt, N] -> out[Nr, N + mem]."""
    N = signal.shape[-1]
    mem = int(delays[-1])
    if taps.ndim == 2:
        out = np.zeros(N + mem, dtype=complex)
        for i, d in enumerate(delays):
            out[d:d + N] += taps[i] * signal
    else:
        _, Nr, Nt, _ = taps.shape
        out = np.zeros((Nr, N + mem), dtype=complex)
        for i, d in enumerate(delays):
            for t in range(Nt):
                out[:, d:d + N] += taps[i, :, t, :] * signal[t]
    return out


def tdl_corrupt_freq(signal, taps, delays, fft_size, carrier_indexes=None):
    """TdlChannel.corrupt_data_in_freq_domain (fading.py:1126-1287), forward direction: block b of
    `block_size` symbols is multiplied by the frequency response of the b-th impulse response
    (taps[..., b], already scaled by sqrt(P)) at the used carriers.
    SISO: signal[N], taps[taps, B] -> out[N].  MIMO: signal[Nt, N], taps[taps, Nr, Nt, B] -> out[Nr, N]."""
    from .ofdm import dense_taps
    car = np.arange(fft_size) if carrier_indexes is None else np.arange(fft_size)[carrier_indexes]
    bs = car.size
    N = signal.shape[-1]
    if N % bs != 0:
        raise ValueError("The num of elements in `signal` must be a multiple of number of sent "
                         "elements per `fft_size`.")
    B = N // bs
    H = np.fft.fft(dense_taps(taps, delays), fft_size, axis=0)[car]          # [bs, (Nr, Nt,) B]
    if taps.ndim == 2:
        out = np.empty(N, dtype=complex)
        for b in range(B):
            out[b * bs:(b + 1) * bs] = H[:, b] * signal[b * bs:(b + 1) * bs]
        return out
    Nr, Nt = taps.shape[1], taps.shape[2]
    out = np.zeros((N, Nr), dtype=complex)
    for b in range(B):
        for t in range(Nt):
            out[b * bs:(b + 1) * bs, :] += H[:, :, t, b] * signal[t, b * bs:(b + 1) * bs, np.newaxis]
    return out.T


def jakes_block_samples(phi, psi, Fd, Ts, t0, num_blocks, fft_size):
    """Fading samples seen by corrupt_data_in_freq_domain: one sample per block, then
    skip_samples_for_next_generation(fft_size - 1) (fading.py:1203-1276, fading_generators.py:525-540).
    Returns (h[*shape, num_blocks], new_t0)."""
    out = []
    for _ in range(num_blocks):
        h, t0 = jakes_samples(phi, psi, Fd, Ts, t0, 1)
        out.append(h)
        t0 += (fft_size - 1) * Ts
    return np.concatenate(out, axis=-1), t0
