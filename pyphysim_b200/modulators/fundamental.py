"""Digital modulators with the API of pyphysim.modulators.fundamental; map / demap run on the GPU.

Constellation tables are built once on the host (they are M <= 256 numbers; SURVEY.md §8a rows
a1/a2); ``modulate`` / ``demodulate`` move the per-symbol work to libb200phy
(``b200phy_map`` / ``b200phy_demap``).  NumPy in -> NumPy out with the reference's dtypes
(int64 indices, complex128 samples); torch CUDA tensors in -> torch tensors out.
"""
import math

import numpy as np

from .. import _device as D
from .. import _lib
from ..util.conversion import binary2gray, dB2Linear, gray2binary
from ..util.misc import level2bits, qfunc

PI = np.pi

__all__ = ['Modulator', 'PSK', 'QPSK', 'BPSK', 'QAM']


class Modulator:
    """Base class (reference: fundamental.py:32-390).  Subclasses call ``setConstellation``."""

    _kind = _lib.MODEM_TABLE

    def __init__(self):
        self._M = 0
        self._K = 0
        self.symbols = np.array([])
        self._tables = D.ModemTables()

    # the device handles are not picklable; SimulationRunner pickles whole runner objects
    # (simulations/runner.py:1836-1846), so drop them like SimulationTracking.__getstate__ does
    def __getstate__(self):
        st = dict(self.__dict__)
        st['_tables'] = None
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._tables = D.ModemTables()

    @property
    def name(self):
        return "{0:d}-{1:s}".format(self._M, self.__class__.__name__)

    @property
    def M(self):
        return self._M

    @property
    def K(self):
        return self._K

    def __repr__(self):
        return "{0} object".format(self.name)

    def setConstellation(self, symbols):
        """fundamental.py:130-145 (K is a float, as in the reference)."""
        M = symbols.size
        self._M = M
        self._K = np.log2(M)
        self.symbols = symbols
        if self._tables is not None:
            self._tables.clear()

    def _native(self, dtype):
        return self._tables.get(self._kind, self.symbols, dtype)

    def modulate(self, inputData):
        """fundamental.py:175-199: gather symbols[inputData]; ValueError on out-of-range."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        scalar = np.isscalar(inputData)
        idx, was_np = D.to_device(np.atleast_1d(inputData) if scalar else inputData, np.int64)
        dtype = _lib.F64
        modem, keep = self._native(dtype)
        out = torch.empty(idx.shape, dtype=_lib.cplx_dtype(dtype), device=idx.device)
        flag = torch.zeros(1, dtype=torch.int32, device=idx.device)
        _lib.check(lib.b200phy_map(dtype, modem, _lib.ptr(idx), idx.numel(), _lib.ptr(out),
                                   _lib.ptr(flag), _lib.cur_stream()))
        if int(flag.item()):
            raise ValueError(self._range_message())
        res = D.from_device(out, was_np)
        return res[0] if scalar else res

    def _range_message(self):
        return "Input data must be between 0 and 2^M"

    def demodulate(self, receivedData):
        """fundamental.py:201-248: index of the closest constellation point."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        dtype = D.dtype_of_samples(receivedData) if hasattr(receivedData, 'dtype') else _lib.F64
        r, was_np = D.to_device(receivedData, D.complex_np(dtype))
        modem, keep = self._native(dtype)
        out = torch.empty(r.shape, dtype=torch.int64, device=r.device)
        _lib.check(lib.b200phy_demap(dtype, modem, _lib.ptr(r), r.numel(), _lib.ptr(out),
                                     _lib.cur_stream()))
        return D.from_device(out, was_np)

    # ---- theory curves (host scalars; fundamental.py:250-390) ---------------------------------
    def calcTheoreticalSER(self, SNR):  # pragma: no cover
        raise NotImplementedError("calcTheoreticalSER: Not implemented")

    def calcTheoreticalBER(self, SNR):  # pragma: no cover
        raise NotImplementedError("calcTheoreticalBER: Not implemented")

    def calcTheoreticalPER(self, SNR, packet_length):
        """fundamental.py:293-320."""
        ber = self.calcTheoreticalBER(SNR)
        return 1 - ((1 - ber) ** packet_length)

    def calcTheoreticalSpectralEfficiency(self, SNR, packet_length=None):
        """fundamental.py:322-390."""
        if packet_length is None:
            return self.K
        per = self.calcTheoreticalPER(SNR, packet_length)
        return self.K * (1 - per)


def _qfunc_array(x):
    return np.vectorize(qfunc)(x) if isinstance(x, np.ndarray) else qfunc(x)


class PSK(Modulator):
    """fundamental.py:393-507."""

    def __init__(self, M, phaseOffset=0):
        super().__init__()
        assert 2 ** math.log(M, 2) == M
        symbols = self._createConstellation(M, phaseOffset)
        symbols = symbols[gray2binary(np.arange(0, M))]         # Gray mapping (fundamental.py:417)
        self.setConstellation(symbols)

    @staticmethod
    def _createConstellation(M, phaseOffset):
        """fundamental.py:420-448 (components below 1e-15 are snapped to zero)."""
        ang = np.arange(0, M) * (2.0 * PI / M) + phaseOffset
        pts = np.stack([np.cos(ang), np.sin(ang)])
        pts[np.abs(pts) < 1e-15] = 0
        return pts[0] + 1j * pts[1]

    def setPhaseOffset(self, phaseOffset):
        """fundamental.py:450-459 — like the reference this rebuilds WITHOUT the Gray reorder."""
        self.setConstellation(self._createConstellation(self._M, phaseOffset))

    def calcTheoreticalSER(self, SNR):
        """fundamental.py:462-483."""
        snr = dB2Linear(SNR)
        return 2. * _qfunc_array(np.sqrt(2. * snr) * math.sin(PI / self._M))

    def calcTheoreticalBER(self, SNR):
        """fundamental.py:485-505."""
        return self.calcTheoreticalSER(SNR) / level2bits(self._M)


class QPSK(PSK):
    """fundamental.py:510-531."""

    def __init__(self):
        super().__init__(4, PI / 4.)
        self._kind = _lib.MODEM_QPSK          # Gray QPSK: the device demaps with a quadrant slicer

    def setConstellation(self, symbols):
        self._kind = _lib.MODEM_TABLE         # any other table (setPhaseOffset drops the Gray order): full search
        super().setConstellation(symbols)

    def __repr__(self):
        return "QPSK object"


class BPSK(Modulator):
    """fundamental.py:534-647."""

    _kind = _lib.MODEM_BPSK

    def __init__(self):
        super().__init__()
        self.setConstellation(np.array([1, -1]))

    @property
    def name(self):
        return "{0:s}".format(self.__class__.__name__)

    def _range_message(self):
        return "Input data can only contains '0's and '1's"

    def modulate(self, inputData):
        """fundamental.py:605-630: 1 - 2*inputData as integers."""
        out = super().modulate(inputData)
        if D.is_torch(out):
            return out
        return np.real(out).astype(np.asarray(inputData).dtype if np.asarray(inputData).dtype.kind in 'iu'
                                   else np.int64)

    def calcTheoreticalSER(self, SNR):
        """fundamental.py:567-588."""
        snr = dB2Linear(SNR)
        return _qfunc_array(np.sqrt(2 * snr))

    def calcTheoreticalBER(self, SNR):
        """fundamental.py:590-603."""
        return self.calcTheoreticalSER(SNR)


class QAM(Modulator):
    """Square Gray-mapped QAM (fundamental.py:656-860)."""

    _kind = _lib.MODEM_QAM

    def __init__(self, M):
        super().__init__()
        power = math.log(M, 2)
        if (power % 2 != 0) or (2 ** power != M):
            raise ValueError("M must be a square power of 2")
        symbols = self._createConstellation(M)
        L = int(round(math.sqrt(M)))
        self.setConstellation(symbols[self._calculateGrayMappingIndexQAM(L)])

    @staticmethod
    def _createConstellation(M):
        """fundamental.py:689-716: raw index ii*L+jj -> (-(L-1)+2jj) + j((L-1)-2ii), unit energy."""
        L = int(round(math.sqrt(M)))
        ii, jj = np.divmod(np.arange(M), L)
        pts = (2.0 * jj - (L - 1)) + 1j * ((L - 1) - 2.0 * ii)
        return pts / math.sqrt((M - 1) * 2.0 / 3.0)

    @staticmethod
    def _calculateGrayMappingIndexQAM(L):
        """fundamental.py:718-777: g[r*L+c] = (gray(r) << log2 L) + gray(c)."""
        g = binary2gray(np.arange(0, L, dtype=int))
        half_bits = level2bits(L ** 2) // 2
        return ((g[:, None] << half_bits) + g[None, :]).reshape(L ** 2)

    def _calcTheoreticalSingleCarrierErrorRate(self, SNR):
        """fundamental.py:779-806."""
        snr = dB2Linear(SNR)
        sqrtM = np.sqrt(self._M)
        return 2. * (1. - (1. / sqrtM)) * _qfunc_array(np.sqrt(snr * 3. / (self._M - 1.)))

    def calcTheoreticalSER(self, SNR):
        """fundamental.py:808-829."""
        Psc = self._calcTheoreticalSingleCarrierErrorRate(SNR)
        return 1. - (1. - Psc) ** 2

    def calcTheoreticalBER(self, SNR):
        """fundamental.py:831-860."""
        k = level2bits(self._M)
        return (2. * self._calcTheoreticalSingleCarrierErrorRate(SNR)) / k
