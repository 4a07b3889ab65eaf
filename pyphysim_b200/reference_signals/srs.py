"""Sounding reference signal sequences with the API of pyphysim/reference_signals/srs.py."""
import math

import numpy as np

from .zadoffchu import get_shifted_root_seq

__all__ = ['get_srs_seq', 'SrsUeSequence', 'UeSequence']


def get_srs_seq(root_seq, n_cs):
    """srs.py:23-48: cyclic shift with denominator 8."""
    return get_shifted_root_seq(root_seq, n_cs, 8)


class UeSequence:
    """srs.py:51-262: a user's sequence; ``normalize`` divides by the norm of the sequence (of its first
    cover-code row when there is a cover-code axis)."""

    def __init__(self, root_seq, n_cs, user_seq_array, normalize=False):
        self._n_cs = n_cs
        self._root_index = root_seq.index
        self._normalized = normalize
        if normalize is True:
            first = user_seq_array if user_seq_array.ndim == 1 else user_seq_array[0]
            user_seq_array = user_seq_array / np.linalg.norm(first)
        self._user_seq_array = user_seq_array

    @property
    def normalized(self):
        return self._normalized

    @property
    def size(self):
        return int(self.seq_array().size)

    @property
    def shape(self):
        return self.seq_array().shape

    def seq_array(self):
        return self._user_seq_array

    def __getitem__(self, val):
        return self.seq_array()[val]

    def __repr__(self):
        return "<{0}(root_index={1}, n_cs={2})>".format(self.__class__.__name__, self._root_index, self._n_cs)

    def __add__(self, other):
        return self.seq_array() + other

    __radd__ = __add__

    def __mul__(self, other):
        return self.seq_array() * other

    __rmul__ = __mul__

    def conjugate(self):
        return self.seq_array().conj()

    conj = conjugate


def _user_rows(root_seq, n_cs, denominator, cover_code, normalize):
    """Rows of a user's sequence from the generator kernel, normalisation folded into the kernel's scale:
    every entry has unit modulus, so the norm of a row is sqrt(size) * |cover entry|."""
    assert abs(n_cs) >= 0
    assert abs(n_cs) < denominator
    if cover_code is None:
        scale = 1.0 / math.sqrt(root_seq.size) if normalize else 1.0
        return root_seq._shifted(n_cs, denominator, scale)
    cc = np.asarray(cover_code)
    norm = math.sqrt(root_seq.size) * abs(complex(cc[0])) if normalize else 1.0
    return np.stack([root_seq._shifted(n_cs, denominator, complex(c) / norm) for c in cc])


class SrsUeSequence(UeSequence):
    """srs.py:265-286."""

    def __init__(self, root_seq, n_cs, normalize=False):
        rows = _user_rows(root_seq, n_cs, 8, None, normalize)
        super().__init__(root_seq, n_cs, rows, normalize=False)
        self._normalized = normalize
