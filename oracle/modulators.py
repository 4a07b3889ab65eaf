"""Constellations, map / min-distance demap, error counting — NumPy restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference: ``pyphysim/modulators/
fundamental.py``, ``pyphysim/util/{misc,conversion}.py``.
"""
import math

import numpy as np


# ---- util/conversion.py ---------------------------------------------------
def dB2Linear(v):
    """util/conversion.py:139-158: pow(10, v/10)."""
    return pow(10, v / 10.0)


def linear2dB(v):
    """util/conversion.py:161-180."""
    return 10.0 * np.log10(v)


def binary2gray(num):
    """util/conversion.py:229-249: (num >> 1) ^ num."""
    return (num >> 1) ^ num


def gray2binary(num):
    """util/conversion.py:252-279: folds shifts 8,4,2,1 only (<= 16-bit values)."""
    t = num ^ (num >> 8)
    t = t ^ (t >> 4)
    t = t ^ (t >> 2)
    t = t ^ (t >> 1)
    return t


# ---- util/misc.py ---------------------------------------------------------
def int2bits(n):
    """util/misc.py:417-446."""
    if n < 0:
        raise ValueError("int2bits: n must be greater then zero")
    if n == 0:
        return 1
    bits = 0
    while n:
        n >>= 1
        bits += 1
    return bits


def level2bits(n):
    """util/misc.py:392-414."""
    if n < 1:
        raise ValueError("level2bits: n must be greater then one")
    return int2bits(n - 1)


def count_bits(n):
    """util/misc.py:449-476 (numba ufunc in the reference): popcount of non-negative ints."""
    n = np.asarray(n, dtype=np.uint64)
    out = np.zeros(n.shape, dtype=np.int64)
    while np.any(n):
        out += (n & np.uint64(1)).astype(np.int64)
        n = n >> np.uint64(1)
    return out


def count_bit_errors(first, second, axis=None):
    """util/misc.py:519-566: sum(popcount(first ^ second))."""
    return np.sum(count_bits(np.bitwise_xor(first, second)), axis)


# ---- constellations -------------------------------------------------------
def qam_constellation(M):
    """QAM.__init__/_createConstellation/_calculateGrayMappingIndexQAM
    (fundamental.py:659-687, 689-716, 718-777)."""
    power = math.log(M, 2)
    if (power % 2 != 0) or (2 ** power != M):
        raise ValueError("M must be a square power of 2")
    L = int(round(math.sqrt(M)))
    raw = np.empty(M, dtype=complex)
    for jj in range(L):
        for ii in range(L):
            raw[ii * L + jj] = complex(-(L - 1) + jj * 2, (L - 1) - ii * 2)
    raw = raw / math.sqrt((M - 1) * 2.0 / 3.0)
    col = binary2gray(np.arange(0, L, dtype=int))
    idx = (np.tile(col.reshape(L, 1), (1, L)) << (level2bits(L ** 2) // 2)) \
        + np.tile(col, (L, 1))
    return raw[idx.reshape(L ** 2)]


def psk_raw(M, phase_offset):
    """PSK._createConstellation (fundamental.py:420-448): snaps |x|<1e-15 to 0."""
    phases = 2.0 * np.pi / M * np.arange(0, M) + phase_offset
    re = np.cos(phases)
    im = np.sin(phases)
    re[abs(re) < 1e-15] = 0
    im[abs(im) < 1e-15] = 0
    return re + 1j * im


def psk_constellation(M, phase_offset=0.0):
    """PSK.__init__ (fundamental.py:396-419): raw[gray2binary(0..M-1)]."""
    assert 2 ** math.log(M, 2) == M
    return psk_raw(M, phase_offset)[gray2binary(np.arange(0, M))]


def qpsk_constellation():
    """QPSK (fundamental.py:510-514) = PSK(4, pi/4)."""
    return psk_constellation(4, np.pi / 4.0)


def bpsk_constellation():
    """BPSK (fundamental.py:534-541): integer table [1, -1]."""
    return np.array([1, -1])


# ---- map / demap ----------------------------------------------------------
def modulate(symbols, idx):
    """Modulator.modulate (fundamental.py:175-199)."""
    try:
        return symbols[idx]
    except IndexError:
        raise ValueError("Input data must be between 0 and 2^M")


def demodulate(symbols, received, chunk=1 << 16):
    """Modulator.demodulate (fundamental.py:241-248): argmin_m |s_m - r| with the
    M x N broadcast, np.abs (hypot) and first-minimum-wins; evaluated in column
    chunks so the temp stays small (same arithmetic per column)."""
    shape = received.shape
    flat = received.flatten()
    out = np.empty(flat.size, dtype=np.int64)
    const = np.reshape(symbols, [symbols.size, 1])
    for s in range(0, flat.size, chunk):
        out[s:s + chunk] = np.abs(const - flat[s:s + chunk]).argmin(axis=0)
    return out.reshape(shape)


def bpsk_modulate(idx):
    """BPSK.modulate (fundamental.py:605-630)."""
    if np.any(idx > 1):
        raise ValueError("Input data can only contains '0's and '1's")
    return 1 - 2 * idx


def bpsk_demodulate(received):
    """BPSK.demodulate (fundamental.py:632-647): (r < 0); NumPy orders complex
    values lexicographically (real part, then imaginary part)."""
    return (received < 0).astype(int)
