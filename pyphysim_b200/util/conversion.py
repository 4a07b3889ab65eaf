"""Host-side scalar conversions with the reference's names and semantics
(pyphysim/util/conversion.py).  These are one-off scalar/table computations ("host,
once"); nothing here touches per-sample data."""
import numpy as np


def dB2Linear(valueIndB):
    """util/conversion.py:139-158."""
    return pow(10, valueIndB / 10.0)


def linear2dB(valueInLinear):
    """util/conversion.py:161-180."""
    return 10.0 * np.log10(valueInLinear)


def dBm2Linear(valueIndBm):
    """util/conversion.py:183-203."""
    return dB2Linear(valueIndBm) / 1000.


def linear2dBm(valueInLinear):
    """util/conversion.py:206-226."""
    return linear2dB(valueInLinear * 1000.)


def binary2gray(num):
    """util/conversion.py:229-249."""
    return (num >> 1) ^ num


def gray2binary(num):
    """util/conversion.py:252-279 (valid up to 16-bit values, like the reference)."""
    t = num ^ (num >> 8)
    t ^= t >> 4
    t ^= t >> 2
    t ^= t >> 1
    return t


def SNR_dB_to_EbN0_dB(SNR, bits_per_symb):
    """util/conversion.py:282-301."""
    return SNR - 10 * np.log10(bits_per_symb)


def EbN0_dB_to_SNR_dB(EbN0, bits_per_symb):
    """util/conversion.py:304-323."""
    return EbN0 + 10 * np.log10(bits_per_symb)
