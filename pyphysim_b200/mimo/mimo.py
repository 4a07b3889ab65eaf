"""MIMO schemes with the API of pyphysim.mimo.mimo: Blast (ZF / MMSE), Alamouti, and (SURVEY.md §8f row
next-3) MRT, MRC, SVDMimo, GMDMimo with the post-processing SINR helpers.

Everything per SYMBOL runs on the GPU: ``b200phy_blast_decode``, ``b200phy_alamouti_encode/decode`` and,
for the precoded schemes, ``b200phy_mat_apply`` (W.X, G_H.Y); the decompositions of the channel matrix
itself run on the GPU too (``b200phy_svd``, ``b200phy_gmd``).  The channel is a tiny host array
(Nr x Nt <= 4 x 4); what stays in NumPy is scalar bookkeeping on it (diag(1/S), the SINR ratio).

SVD gauge: ``b200phy_svd`` fixes the phase of every singular pair (largest entry of v_i real positive)
where numpy returns LAPACK's; precoder and receive filter therefore differ from the reference's by one
unit phase per stream, which cancels in G_H H W (oracle/mimo.py ``svd_canonical``, tests).
"""
import math
import warnings

import numpy as np

from .. import _device as D
from .. import _lib

__all__ = ['MimoBase', 'MisoBase', 'Blast', 'Alamouti', 'MRT', 'MRC', 'SVDMimo', 'GMDMimo',
           'calc_post_processing_SINRs', 'calc_post_processing_linear_SINRs']


def calc_post_processing_linear_SINRs(channel, W, G_H, noise_var=None):
    """mimo.py:63-114: per stream |diag|^2 / (|row sum - diag|^2 + noise_var ||row of G_H||^2) of the
    equivalent channel G_H H W.  Scalar bookkeeping on matrices of at most 4 x 4: host NumPy."""
    if noise_var is None:
        noise_var = 0.0
    Heq = np.atleast_2d(np.dot(G_H, np.dot(channel, W)))
    diag = np.diag(Heq)
    leak = np.sum(Heq, axis=1) - diag
    if isinstance(G_H, np.ndarray):
        gain = np.linalg.norm(np.atleast_2d(G_H), axis=1) ** 2
    else:
        gain = abs(G_H) ** 2
    return np.abs(diag) ** 2 / (np.abs(leak) ** 2 + noise_var * gain)


def calc_post_processing_SINRs(channel, W, G_H, noise_var=None):
    """mimo.py:26-60: the same in dB."""
    return 10.0 * np.log10(calc_post_processing_linear_SINRs(channel, W, G_H, noise_var))


def _mat_apply(A, X):
    """A (host, small) applied to the columns of X (host or device) on the GPU."""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    x, was_np = D.to_device(X, np.complex128)
    if x.dim() == 1:
        x = x.reshape(1, -1)
    A = np.ascontiguousarray(np.atleast_2d(np.asarray(A, dtype=np.complex128)))
    rows, cols = A.shape
    if cols != x.shape[0]:
        raise ValueError("shapes (%d,%d) and (%d,%d) not aligned" % (rows, cols, x.shape[0], x.shape[1]))
    a, _ = D.to_device(A, np.complex128)
    n = x.shape[1]
    y = torch.empty((rows, n), dtype=torch.complex128, device='cuda')
    _lib.check(lib.b200phy_mat_apply(_lib.F64, _lib.ptr(a), rows, cols, _lib.ptr(x.contiguous()), n,
                                     _lib.ptr(y), _lib.cur_stream()))
    return y, was_np


def _svd(channel):
    """Thin gauge-fixed SVD of one channel on the GPU -> host (U [Nr, Nt], S [Nt], V [Nt, Nt])."""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    H = np.ascontiguousarray(np.asarray(channel, dtype=np.complex128))
    Nr, Nt = H.shape
    if Nt > Nr:
        raise ValueError("the decomposition needs Nt <= Nr (got %d x %d)" % (Nr, Nt))
    h, _ = D.to_device(H, np.complex128)
    U = torch.empty((Nr, Nt), dtype=torch.complex128, device='cuda')
    S = torch.empty(Nt, dtype=torch.float64, device='cuda')
    V = torch.empty((Nt, Nt), dtype=torch.complex128, device='cuda')
    _lib.check(lib.b200phy_svd(_lib.ptr(h), 1, Nr, Nt, _lib.ptr(U), _lib.ptr(S), _lib.ptr(V), _lib.cur_stream()))
    return U, S, V


def _gmd_dev(U, S, V):
    lib = _lib.load()
    torch = _lib.torch_cuda()
    Nr, Nt = U.shape
    Q, R, P = torch.empty_like(U), torch.empty((Nt, Nt), dtype=torch.float64, device='cuda'), torch.empty_like(V)
    _lib.check(lib.b200phy_gmd(_lib.ptr(U), _lib.ptr(S), _lib.ptr(V), 1, Nr, Nt, _lib.ptr(Q), _lib.ptr(R),
                               _lib.ptr(P), _lib.cur_stream()))
    return Q, R, P


class MimoBase:
    """reference: mimo/mimo.py:117-462 (the parts Blast and Alamouti need)."""

    def __init__(self, channel=None):
        self._channel = None
        if channel is not None:
            self.set_channel_matrix(channel)

    def set_channel_matrix(self, channel):
        """mimo.py:171-190 (a 1-D array is a single-receive-antenna channel)."""
        channel = np.asarray(channel)
        if channel.ndim == 1:
            channel = channel[np.newaxis, :]
        self._channel = channel.astype(complex)

    @property
    def Nt(self):
        return self._channel.shape[1]

    @property
    def Nr(self):
        return self._channel.shape[0]

    def getNumberOfLayers(self):  # pragma: no cover
        raise NotImplementedError("getNumberOfLayers still needs to be implemented in the subclass")

    @staticmethod
    def _calc_precoder(channel):  # pragma: no cover
        raise NotImplementedError('_calc_precoder still needs to be implemented')

    @staticmethod
    def _calc_receive_filter(channel, noise_var=None):  # pragma: no cover
        raise NotImplementedError('_calc_receive_filter still needs to be implemented')

    def calc_linear_SINRs(self, noise_var):
        """mimo.py:311-333.  As in the reference this returns the post-processing SINRs in dB (it calls
        calc_post_processing_SINRs, and the reference's own tests expect dB here)."""
        W = self._calc_precoder(self._channel)
        G_H = self._calc_receive_filter(self._channel, noise_var)
        return calc_post_processing_SINRs(self._channel, W, G_H, noise_var)

    def calc_SINRs(self, noise_var):
        """mimo.py:335-352: linear2dB of the above."""
        return 10.0 * np.log10(self.calc_linear_SINRs(noise_var))

    def encode(self, transmit_data):  # pragma: no cover
        raise NotImplementedError("encode still needs to be implemented in the subclass")

    def decode(self, received_data):  # pragma: no cover
        raise NotImplementedError("decode still needs to be implemented in the subclass")

    def _channel_dev(self):
        t, _ = D.to_device(self._channel, np.complex128)
        return t


class Blast(MimoBase):
    """Spatial multiplexing with a linear ZF / MMSE receiver (mimo.py:465-660)."""

    def __init__(self, channel=None):
        super().__init__(channel)
        self._noise_var = 0.0

    def set_channel_matrix(self, channel):
        Nr, Nt = np.asarray(channel).shape
        if Nt > Nr:
            warnings.warn("The number of transmit antennas for {0} should not be greater than the "
                          "number of receive antennas.".format(self.__class__.__name__))
        super().set_channel_matrix(channel)

    def getNumberOfLayers(self):
        return self.Nt

    def set_noise_var(self, noise_var):
        """mimo.py:531-553: None or 0 -> zero forcing, > 0 -> MMSE."""
        if noise_var is None:
            self._noise_var = 0.0
        elif noise_var >= 0.0:
            self._noise_var = noise_var
        else:
            raise ValueError('Noise variance must be a non-negative value.')

    @staticmethod
    def _calc_precoder(channel):
        """mimo.py:555-588: identity / sqrt(Nt)."""
        Nt = np.asarray(channel).shape[1]
        return np.eye(Nt) / math.sqrt(Nt)

    @staticmethod
    def _calc_receive_filter(channel, noise_var=None):
        """mimo.py:590-607: sqrt(Nt) x (MMSE filter if noise_var > 0 else pseudo-inverse), computed by
        ``b200phy_blast_decode`` applied to the identity."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        H = np.ascontiguousarray(np.atleast_2d(np.asarray(channel, dtype=np.complex128)))
        Nr, Nt = H.shape
        h, _ = D.to_device(H, np.complex128)
        eye = torch.eye(Nr, dtype=torch.complex128, device='cuda')
        out = torch.empty(Nt * Nr, dtype=torch.complex128, device='cuda')
        _lib.check(lib.b200phy_blast_decode(_lib.F64, _lib.ptr(h), _lib.ptr(eye), 1, Nr, Nt, Nr,
                                            float(noise_var or 0.0), _lib.ptr(out), _lib.cur_stream()))
        return out.reshape(Nr, Nt).t().cpu().numpy()               # decode output is order='F'

    def encode(self, transmit_data):
        """mimo.py:609-640: symbol k goes to antenna k mod Nt, power split 1/sqrt(Nt)."""
        from ..channels.fading import _scale_rows
        x, was_np = D.to_device(transmit_data, np.complex128)
        n_streams = self.getNumberOfLayers()
        if x.numel() % n_streams != 0:
            raise ValueError("Input array number of elements must be a multiple of the"
                             " number of transmit antennas")
        enc = x.reshape(-1, n_streams).t().contiguous()            # reshape(nStreams, -1, order='F')
        _scale_rows(enc, [1.0 / math.sqrt(self.Nt)])
        return D.from_device(enc, was_np)

    def decode(self, received_data):
        """mimo.py:642-660."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        y, was_np = D.to_device(received_data, np.complex128)
        if y.dim() == 1:
            y = y.reshape(-1, 1)
        Nr, T = y.shape
        if Nr != self.Nr:
            raise ValueError("received_data must have %d rows" % self.Nr)
        H = self._channel_dev()
        out = torch.empty(self.Nt * T, dtype=torch.complex128, device='cuda')
        _lib.check(lib.b200phy_blast_decode(_lib.F64, _lib.ptr(H), _lib.ptr(y.contiguous()), 1, self.Nr,
                                            self.Nt, T, float(self._noise_var), _lib.ptr(out),
                                            _lib.cur_stream()))
        return D.from_device(out, was_np)


class Alamouti(MimoBase):
    """Alamouti space-time block code, 2 transmit antennas (mimo.py:1073-1287)."""

    def set_channel_matrix(self, channel):
        channel = np.asarray(channel)
        if channel.ndim == 1:
            super().set_channel_matrix(channel[np.newaxis, :])
        else:
            _, Nt = channel.shape
            if Nt != 2:
                raise ValueError("The number of transmit antennas must be equal to 2 for the "
                                 "{0} scheme".format(self.__class__.__name__))
            super().set_channel_matrix(channel)

    def getNumberOfLayers(self):
        return 1

    def calc_linear_SINRs(self, noise_var):
        """mimo.py:1133-1164: ||H||_F^2 / noise_var (host scalar)."""
        return np.linalg.norm(self._channel, 'fro') ** 2 / noise_var

    def calc_SINRs(self, noise_var):
        return 10.0 * np.log10(self.calc_linear_SINRs(noise_var))

    def encode(self, transmit_data):
        """mimo.py:1166-1214."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        s, was_np = D.to_device(transmit_data, np.complex128)
        s = s.reshape(-1)
        T = s.numel()
        x = torch.empty((2, T), dtype=torch.complex128, device='cuda')
        _lib.check(lib.b200phy_alamouti_encode(_lib.F64, _lib.ptr(s), 1, T, _lib.ptr(x), _lib.cur_stream()))
        return D.from_device(x, was_np)

    def decode(self, received_data):
        """mimo.py:1216-1287."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        y, was_np = D.to_device(received_data, np.complex128)
        if y.dim() == 1:
            y = y.reshape(1, -1)
        Nr, T = y.shape
        H = self._channel_dev()
        out = torch.empty(T, dtype=torch.complex128, device='cuda')
        _lib.check(lib.b200phy_alamouti_decode(_lib.F64, _lib.ptr(H), _lib.ptr(y.contiguous()), 1, Nr, T,
                                               _lib.ptr(out), _lib.cur_stream()))
        return D.from_device(out, was_np)


class MisoBase(MimoBase):
    """Schemes with one receive antenna and one stream (mimo.py:355-462)."""

    def __init__(self, channel=None):
        super().__init__(channel=None)
        if channel is not None:
            self.set_channel_matrix(channel)

    def set_channel_matrix(self, channel):
        channel = np.asarray(channel)
        if channel.ndim == 1:
            super().set_channel_matrix(channel[np.newaxis, :])
        else:
            if channel.shape[0] != 1:
                raise ValueError("The MRT scheme is only defined for the "
                                 "scenario with a single receive antenna")
            super().set_channel_matrix(channel)

    def getNumberOfLayers(self):
        return 1


class MRT(MisoBase):
    """Maximum ratio transmission (mimo.py:666-783): co-phase the transmit antennas."""

    @staticmethod
    def _calc_precoder(channel):
        """mimo.py:688-712: exp(-j angle(h))^T / sqrt(Nt)."""
        channel = np.atleast_2d(np.asarray(channel))
        return np.exp(-1j * np.angle(channel)).T / math.sqrt(channel.shape[1])

    @staticmethod
    def _calc_receive_filter(channel, noise_var=None):
        """mimo.py:714-735: the scalar sqrt(Nt) / sum |h|."""
        channel = np.atleast_2d(np.asarray(channel))
        return math.sqrt(channel.shape[1]) / np.sum(np.abs(channel))

    def encode(self, transmit_data):
        """mimo.py:737-761: [Nt, n] = W x."""
        y, was_np = _mat_apply(self._calc_precoder(self._channel), transmit_data)
        return D.from_device(y, was_np)

    def decode(self, received_data):
        """mimo.py:763-783: G_H y, flattened."""
        G = np.array([[self._calc_receive_filter(self._channel)]], dtype=np.complex128)
        y, was_np = _mat_apply(G, received_data)
        return D.from_device(y.reshape(-1), was_np)


class MRC(Blast):
    """Maximum ratio combining (mimo.py:786-826): Blast whose channel may be a 1-D column."""

    def set_channel_matrix(self, channel):
        channel = np.asarray(channel)
        if channel.ndim == 1:
            super().set_channel_matrix(channel[:, np.newaxis])
        else:
            super().set_channel_matrix(channel)


class SVDMimo(Blast):
    """Eigen-beamforming: precode with V, receive with diag(1/S) U^H (mimo.py:829-948)."""

    def set_channel_matrix(self, channel):
        Nr, Nt = np.asarray(channel).shape
        if Nr != Nt:
            # the reference's diag(1/S).dot(U^H) only has matching shapes for a square channel
            raise ValueError("SVDMimo needs a square channel matrix (got %d x %d)" % (Nr, Nt))
        super().set_channel_matrix(channel)

    @staticmethod
    def _calc_precoder(channel):
        """mimo.py:855-874: V / sqrt(Nt) (gauge-fixed SVD, see the module docstring)."""
        _, _, V = _svd(channel)
        return V.cpu().numpy() / math.sqrt(np.asarray(channel).shape[1])

    @staticmethod
    def _calc_receive_filter(channel, noise_var=None):
        """mimo.py:876-898: diag(1/S) U^H sqrt(Nt)."""
        U, S, _ = _svd(channel)
        U, S = U.cpu().numpy(), S.cpu().numpy()
        return np.diag(1.0 / S).dot(U.conj().T) * math.sqrt(np.asarray(channel).shape[1])

    def encode(self, transmit_data):
        """mimo.py:900-928: W . reshape(x, (Nt, -1))."""
        x, was_np = D.to_device(transmit_data, np.complex128)
        if x.numel() % self.Nt != 0:
            raise ValueError("Input array number of elements must be a multiple of the"
                             " number of transmit antennas")
        y, _ = _mat_apply(self._calc_precoder(self._channel), x.reshape(self.Nt, -1))
        return D.from_device(y, was_np)

    def decode(self, received_data):
        """mimo.py:930-948: G_H . y, flattened row by row."""
        y, was_np = _mat_apply(self._calc_receive_filter(self._channel), received_data)
        return D.from_device(y.reshape(-1), was_np)


class GMDMimo(Blast):
    """Geometric mean decomposition precoding (mimo.py:951-1067): every stream sees the same gain."""

    @staticmethod
    def _decompose(channel):
        U, S, V = _svd(channel)
        return _gmd_dev(U, S, V)

    @staticmethod
    def _calc_precoder(channel):
        """mimo.py:974-994: P / sqrt(Nt)."""
        _, _, P = GMDMimo._decompose(channel)
        return P.cpu().numpy() / math.sqrt(np.asarray(channel).shape[1])

    @staticmethod
    def _calc_receive_filter(channel, noise_var=None):
        """mimo.py:996-1019: the Blast filter of the equivalent channel Q R."""
        Q, R, _ = GMDMimo._decompose(channel)
        Heq = Q.cpu().numpy().dot(R.cpu().numpy())
        return Blast._calc_receive_filter(Heq, noise_var)

    def encode(self, transmit_data):
        """mimo.py:1021-1048."""
        x, was_np = D.to_device(transmit_data, np.complex128)
        if x.numel() % self.Nt != 0:
            raise ValueError("Input array number of elements must be a multiple of the"
                             " number of transmit antennas")
        y, _ = _mat_apply(self._calc_precoder(self._channel), x.reshape(self.Nt, -1))
        return D.from_device(y, was_np)

    def decode(self, received_data):
        """mimo.py:1050-1067: Blast decode over the equivalent channel, flattened row by row."""
        y, was_np = _mat_apply(self._calc_receive_filter(self._channel, self._noise_var), received_data)
        return D.from_device(y.reshape(-1), was_np)
