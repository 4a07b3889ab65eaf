"""RootSequence with the API of pyphysim/reference_signals/root_sequence.py:221-496.

Sequences longer than two PRBs are (cyclically extended) Zadoff-Chu sequences; 12- and 24-element sequences come
from the phase tables of 3GPP TS 36.211 (Tables 5.5.1.2-1 / -2), stored here two bits per entry: entry n of
row ``root_index`` is ``2 * ((word >> 2n) & 3) - 3``.  Generation runs on the GPU (``b200phy_refsig_sequence``)."""
import numpy as np

from .zadoffchu import _sequence

__all__ = ['RootSequence']

_PHI12 = (0xcbaf39, 0xc827fa, 0x62040a, 0x7206a9, 0xe6466d, 0x8d6972, 0x8f6c0d, 0xb27254, 0x9e95b2, 0xaa87d2,
          0x70429d, 0xfb8f5b, 0x80a8a2, 0xf1e8cf, 0x9fed18, 0x1ba527, 0x75fe6e, 0x7b0ce8, 0x2508ac, 0x44d6ed,
          0x49baa1, 0x18009d, 0xc8d00a, 0x9e611a, 0x819fba, 0xd4bef2, 0x1d630e, 0x0e6c44, 0x5f3dcd, 0x6cd143)
_PHI24 = (0x1339acece72d, 0x2b836ea7080c, 0xf9469dbfcaf7, 0x44adbb94a3a1, 0x520619dfa415, 0xa031d688b9e8,
          0x1d91a658c35a, 0x94e6daeed17c, 0x2020255cc638, 0xa298a7f347ca, 0xd9f981017709, 0x89c7cf01b83e,
          0x44c7ca725abe, 0xc78a7f9bd157, 0xeb485faecbb0, 0x40506ed18e25, 0x4a226fb29571, 0x5747a7f187de,
          0x2b0a48e876aa, 0x0de4663f71be, 0x63674ec45031, 0x24f5992d99a0, 0x1f8b933046c4, 0xf6cf778fbf55,
          0xd16a977531f6, 0x50a38a16b766, 0xa898cb314ae4, 0x444d0b547af1, 0x5dac22865251, 0xe583b6e7745a)


def _phi_row(size, root_index):
    try:
        w = (_PHI12 if size == 12 else _PHI24)[int(root_index)]
    except IndexError:
        raise KeyError('{0}'.format(root_index))            # the reference indexes a dict of 30 rows
    return [2 * ((w >> (2 * n)) & 3) - 3 for n in range(size)]


def _largest_prime_leq(n):
    """root_sequence.py:289-305 (the reference's table stops at 1009)."""
    n = min(int(n), 1009)
    while n >= 2:
        if all(n % d for d in range(2, int(n ** 0.5) + 1)):
            return n
        n -= 1
    raise IndexError('no prime number lower than or equal to the sequence size')


class RootSequence:
    """root_sequence.py:221-287.  ``root_index``: SRS root sequence index; ``size``: size after cyclic extension;
    ``Nzc``: Zadoff-Chu length (default: the largest prime <= size)."""
    n_sc_PRB = 12

    def __init__(self, root_index, size=None, Nzc=None):
        if size is None and Nzc is None:
            raise AttributeError("Either 'size' or 'Nzc' (or both) must be provided.")
        if size is None:
            size = Nzc
        assert isinstance(size, int)
        if Nzc is None:
            Nzc = self._get_largest_prime_lower_than_number(size)
        if size < Nzc:
            raise AttributeError("If 'size' and Nzc are provided, then size must be greater than Nzc")
        self._root_index = root_index
        self._extended_seq_array = None
        self._phi = None
        if size > 2 * self.n_sc_PRB:
            assert root_index < Nzc
            full = _sequence(Nzc, root_index, size)            # extension = index mod Nzc on the device
            self._seq_array = full[:Nzc]
            if size > Nzc:
                self._extended_seq_array = full
        elif size in (self.n_sc_PRB, 2 * self.n_sc_PRB):
            self._phi = _phi_row(size, root_index)
            self._seq_array = _sequence(size, 0, size, phi=self._phi)
        else:
            raise AttributeError("Invalid root sequence size")

    @staticmethod
    def _get_largest_prime_lower_than_number(seq_size):
        return _largest_prime_leq(seq_size)

    @property
    def Nzc(self):
        return int(self._seq_array.size)

    @property
    def size(self):
        """Size with extension (== Nzc when the sequence is not extended), root_sequence.py:320-344."""
        if self._extended_seq_array is None:
            return self.Nzc
        return int(self._extended_seq_array.size)

    @property
    def index(self):
        return self._root_index

    def seq_array(self):
        """root_sequence.py:358-370."""
        if self._extended_seq_array is None:
            return self._seq_array
        return self._extended_seq_array

    def _shifted(self, n_cs, denominator, scale=1.0):
        """The user sequence of this root (cyclic shift n_cs / denominator, times ``scale``) straight from the
        generator kernel — what SrsUeSequence / DmrsUeSequence are made of."""
        if self._phi is not None:
            return _sequence(self.Nzc, 0, self.size, n_cs, denominator, phi=self._phi, scale=scale)
        return _sequence(self.Nzc, self._root_index, self.size, n_cs, denominator, scale=scale)

    def __add__(self, other):
        return self.seq_array() + other

    __radd__ = __add__

    def __mul__(self, other):
        return self.seq_array() * other

    __rmul__ = __mul__

    def __getitem__(self, val):
        return self.seq_array()[val]

    def conjugate(self):
        return self.seq_array().conj()

    conj = conjugate

    def __repr__(self):
        if self._extended_seq_array is None:
            return "<SrsRootSequence(root_index={0},Nzc={1})>".format(self._root_index, self._seq_array.size)
        return "<SrsRootSequence(root_index={0},size={2},Nzc={1})>".format(
            self._root_index, self._seq_array.size, self._extended_seq_array.size)
