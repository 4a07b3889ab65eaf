"""Demodulation reference signal sequences with the API of pyphysim/reference_signals/dmrs.py."""
import numpy as np

from .srs import UeSequence, _user_rows
from .zadoffchu import get_shifted_root_seq

__all__ = ['get_dmrs_seq', 'DmrsUeSequence']


def get_dmrs_seq(root_seq, n_cs):
    """dmrs.py:19-41: cyclic shift with denominator 12."""
    return get_shifted_root_seq(root_seq, n_cs, 12)


class DmrsUeSequence(UeSequence):
    """dmrs.py:44-115; ``cover_code``: optional orthogonal cover code, becomes the leading axis of the sequence."""

    def __init__(self, root_seq, n_cs, cover_code=None, normalize=False):
        self._occ = cover_code
        if cover_code is not None:
            assert isinstance(self._occ, np.ndarray)
            self._occ.flags.writeable = False
        rows = _user_rows(root_seq, n_cs, 12, cover_code, normalize)
        super().__init__(root_seq, n_cs, rows, normalize=False)
        self._normalized = normalize

    @property
    def cover_code(self):
        return self._occ

    @property
    def size(self):
        """dmrs.py:91-102: number of elements of ONE cover row."""
        if self._occ is None:
            return int(self._user_seq_array.shape[0])
        return int(self._user_seq_array.shape[1])

    def __repr__(self):
        return "<{0}(root_index={1}, n_cs={2}, cover_code={3})>".format(
            self.__class__.__name__, self._root_index, self._n_cs, self._occ)
