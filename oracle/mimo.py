"""Blast (ZF / MMSE) and Alamouti — NumPy restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference: ``pyphysim/mimo/mimo.py``.
"""
import math

import numpy as np


def zf_filter(H):
    """MimoBase._calcZeroForceFilter (mimo.py:264-285): pinv(H)."""
    return np.linalg.pinv(H)


def mmse_filter(H, noise_var):
    """MimoBase._calcMMSEFilter (mimo.py:287-309): solve(H^H H + s2 I, H^H)."""
    Hh = H.conj().T
    return np.linalg.solve(np.dot(Hh, H) + noise_var * np.eye(H.shape[1]), Hh)


def blast_receive_filter(H, noise_var=0.0):
    """Blast._calc_receive_filter (mimo.py:590-607): MMSE iff noise_var > 0, x sqrt(Nt)."""
    G = mmse_filter(H, noise_var) if noise_var > 0 else zf_filter(H)
    return G * math.sqrt(H.shape[1])


def blast_encode(x, Nt):
    """Blast.encode (mimo.py:609-640): symbol k -> antenna k mod Nt, / sqrt(Nt)."""
    if x.size % Nt != 0:
        raise ValueError("Input array number of elements must be a multiple of the"
                         " number of transmit antennas")
    return x.reshape((Nt, -1), order='F') / math.sqrt(Nt)


def blast_decode(y, H, noise_var=0.0):
    """Blast.decode (mimo.py:642-660)."""
    return blast_receive_filter(H, noise_var).dot(y).reshape(-1, order='F')


def alamouti_encode(s):
    """Alamouti.encode/_encode (mimo.py:1166-1214): [[s0, -s1*], [s1, s0*]] / sqrt(2)."""
    Ns = s.size
    out = np.empty((2, Ns), dtype=complex)
    out[0, 0::2] = s[0::2]
    out[0, 1::2] = -np.conj(s[1::2])
    out[1, 0::2] = s[1::2]
    out[1, 1::2] = np.conj(s[0::2])
    return out / math.sqrt(2)


def alamouti_decode(y, H):
    """Alamouti.decode/_decode (mimo.py:1216-1287); H is [Nr, 2]."""
    H = np.atleast_2d(H)
    h0, h1 = H[:, 0], H[:, 1]
    y0, y1 = y[:, 0::2], y[:, 1::2]
    out = np.empty(y.shape[1], dtype=complex)
    out[0::2] = h0.conj() @ y0 + h1 @ y1.conj()
    out[1::2] = h1.conj() @ y0 + (-h0) @ y1.conj()
    out /= np.linalg.norm(H, 'fro') ** 2
    return out * math.sqrt(2)
