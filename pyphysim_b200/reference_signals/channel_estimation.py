"""CAZAC-based channel estimators with the API of pyphysim/reference_signals/channel_estimation.py, evaluated by
``b200phy_cazac_estimate`` (one CTA per received vector).  NumPy in -> NumPy out (complex128, the reference's
dtype); CUDA tensors in -> CUDA tensors out, and a leading batch axis is accepted through ``estimate_batch``."""
import ctypes as C

import numpy as np

from .. import _device as D
from .. import _lib
from .srs import UeSequence

__all__ = ['CazacBasedChannelEstimator', 'CazacBasedWithOCCChannelEstimator']


def _cazac(ref, y, cover, n_cover, batch, Nsc, num_taps_to_keep, mult, scale, dtype):
    """y: device tensor [batch, n_cover, Nsc] (contiguous) -> device tensor [batch, mult * Nsc]."""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    out = torch.empty((batch, mult * Nsc), dtype=_lib.cplx_dtype(dtype), device='cuda')
    cc = None
    if cover is not None:
        flat = np.asarray(cover, dtype=np.complex128).view(np.float64)
        cc = (C.c_double * flat.size)(*flat)
    _lib.check(lib.b200phy_cazac_estimate(dtype, _lib.ptr(ref), _lib.ptr(y), cc, n_cover, batch, Nsc,
                                          int(num_taps_to_keep), int(mult), float(scale), _lib.ptr(out),
                                          _lib.cur_stream()))
    return out


class CazacBasedChannelEstimator:
    """channel_estimation.py:15-131.  ``ue_ref_seq``: SrsUeSequence / DmrsUeSequence or a plain array;
    ``size_multiplier``: 2 for the SRS comb pattern, 1 otherwise."""

    def __init__(self, ue_ref_seq, size_multiplier=2):
        if isinstance(ue_ref_seq, UeSequence):
            self._normalized_ref_seq = ue_ref_seq.normalized
            ue_ref_seq = ue_ref_seq.seq_array()
        else:
            self._normalized_ref_seq = False
        self._ue_ref_sequence = ue_ref_seq
        self._size_multiplier = size_multiplier
        self._dev = {}

    @property
    def ue_ref_seq(self):
        return self._ue_ref_sequence

    def _ref_on_device(self, dtype):
        torch = _lib.torch_cuda()
        key = (dtype, torch.cuda.current_device())
        if key not in self._dev:
            self._dev[key], _ = D.to_device(np.asarray(self._ue_ref_sequence).reshape(-1), D.complex_np(dtype))
        return self._dev[key]

    def __getstate__(self):                                    # device handles do not pickle
        st = dict(self.__dict__)
        st['_dev'] = {}
        return st

    def _run(self, received_signal, num_taps_to_keep, cover, n_cover):
        r = np.asarray(self.ue_ref_seq)
        Nsc = r.size
        dtype = D.dtype_of_samples(received_signal) if D.is_torch(received_signal) else _lib.F64
        y, was_np = D.to_device(received_signal, D.complex_np(dtype))
        lead = tuple(y.shape[:-2]) if n_cover > 1 else tuple(y.shape[:-1])
        batch = int(np.prod(lead)) if lead else 1
        scale = float(Nsc) if self._normalized_ref_seq is True else 1.0
        out = _cazac(self._ref_on_device(dtype), y.reshape(batch, n_cover, Nsc), cover, n_cover, batch, Nsc,
                     num_taps_to_keep, self._size_multiplier, scale, dtype)
        return D.from_device(out.reshape(lead + (self._size_multiplier * Nsc,)), was_np)

    def estimate_channel_freq_domain(self, received_signal, num_taps_to_keep):
        """channel_estimation.py:69-131: received_signal [Nsc] or [Nr, Nsc] -> [mult*Nsc] or [Nr, mult*Nsc]."""
        if received_signal.ndim not in (1, 2):
            raise ValueError("received_signal must have either one dimension (one receive antenna) or two "
                             "dimensions (first dimension being the receive antenna dimension).")
        return self._run(received_signal, num_taps_to_keep, None, 1)

    def estimate_batch(self, received_signal, num_taps_to_keep):
        """Any number of leading axes (realizations, antennas): one kernel launch for all of them."""
        return self._run(received_signal, num_taps_to_keep, None, 1)


class CazacBasedWithOCCChannelEstimator(CazacBasedChannelEstimator):
    """channel_estimation.py:134-251: the cover code is removed by averaging over the cover axis."""

    def __init__(self, ue_ref_seq):
        cover_code = ue_ref_seq.cover_code
        reference_seq = ue_ref_seq.seq_array()[0] * cover_code[0]
        super().__init__(reference_seq, size_multiplier=1)
        self._cover_code = cover_code
        self._normalized_ref_seq = ue_ref_seq.normalized

    @property
    def cover_code(self):
        return self._cover_code

    def estimate_channel_freq_domain(self, received_signal, num_taps_to_keep, extra_dimension=True):
        """:163-251.  extra_dimension=False: the cover axis is folded into the last one ([n_cc * Nsc])."""
        n_cc = int(np.asarray(self.cover_code).size)
        r = received_signal
        if extra_dimension is False:
            if r.ndim == 1:
                r = r.reshape(n_cc, -1)
            elif r.ndim == 2:
                r = r.reshape(r.shape[0], n_cc, -1)
            else:
                raise RuntimeError('Invalid dimension for received_signal: {0}'.format(r.ndim))
        if r.ndim not in (2, 3):
            raise RuntimeError('Invalid dimension for received_signal: {0}'.format(r.ndim))
        return self._run(r, num_taps_to_keep, np.asarray(self.cover_code), n_cc)
