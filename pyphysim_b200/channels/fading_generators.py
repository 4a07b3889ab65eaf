"""Fading sample generators with the API of pyphysim.channels.fading_generators.

The sample arithmetic runs on the GPU (``b200phy_jakes`` for the Jakes sum of sinusoids, the Philox
stream for Rayleigh samples); the generators keep their samples as device tensors and hand out NumPy
arrays (complex128, the reference's dtype) on request.
"""
import math

import numpy as np

from .. import _lib
from ..util.misc import randn_c

__all__ = ['FadingSampleGenerator', 'RayleighSampleGenerator', 'JakesSampleGenerator']


class FadingSampleGenerator:
    """Base class (reference: fading_generators.py:104-205)."""

    def __init__(self, shape=None):
        self._shape = None
        self._set_shape(shape)
        self._samples_dev = None       # device tensor, complex128
        self._samples_np = None

    def _set_shape(self, shape):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        self._shape = None if shape is None else tuple(int(v) for v in shape)

    @property
    def shape(self):
        return self._shape

    @shape.setter
    def shape(self, value):
        self._set_shape(value)

    def _store(self, dev):
        self._samples_dev = dev
        self._samples_np = None

    def get_samples(self):
        """The last generated samples as a NumPy array (fading_generators.py:152-162)."""
        if self._samples_np is None:
            out = self._samples_dev.cpu().numpy()
            self._samples_np = complex(out) if out.ndim == 0 else out
        return self._samples_np

    def _device_samples(self):
        return self._samples_dev

    def __getstate__(self):            # device tensors do not cross process boundaries
        st = dict(self.__dict__)
        if st.get('_samples_dev') is not None:
            st['_samples_np'] = self.get_samples()
        st['_samples_dev'] = None
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        if self._samples_np is not None and self._samples_dev is None:
            import torch
            if torch.cuda.is_available():
                self._samples_dev = torch.from_numpy(np.asarray(self._samples_np, dtype=complex)).cuda()

    def generate_more_samples(self, num_samples=None):  # pragma: no cover
        raise NotImplementedError("Implement in a subclass")

    def skip_samples_for_next_generation(self, num_samples):  # pragma: no cover
        raise NotImplementedError("Implement in a subclass")

    def get_similar_fading_generator(self):  # pragma: no cover
        raise NotImplementedError("Implement in a subclass")


class RayleighSampleGenerator(FadingSampleGenerator):
    """i.i.d. CN(0,1) samples (fading_generators.py:208-282)."""

    def __init__(self, shape=None):
        super().__init__(shape)
        self.generate_more_samples()

    def generate_more_samples(self, num_samples=None):
        shape = self.shape
        if num_samples is None:
            dims = () if shape is None else shape
        elif shape is None:
            dims = (int(num_samples),)
        else:
            dims = tuple(shape) + (int(num_samples),)
        self._store(randn_c(*dims, device_out=True))

    def skip_samples_for_next_generation(self, num_samples):
        """Samples are independent: nothing to skip (fading_generators.py:251-268)."""

    def get_similar_fading_generator(self):
        return RayleighSampleGenerator(self._shape)


class JakesSampleGenerator(FadingSampleGenerator):
    """h(t) = L^-1/2 sum_l exp(j(2 pi Fd cos(phi_l) t + psi_l))  (fading_generators.py:289-553)."""

    def __init__(self, Fd=100, Ts=1e-3, L=8, shape=None, RS=None):
        super().__init__(shape)
        self._Fd = Fd
        self._Ts = Ts
        self._L = L
        self._phi_l = None
        self._psi_l = None
        self.RS = np.random if RS is None else RS
        self._current_time = 0.0
        self._set_phi_and_psi_according_to_shape()
        self.generate_more_samples()       # like the reference: the clock starts at Ts afterwards

    @property
    def shape(self):
        return self._shape

    @shape.setter
    def shape(self, new_shape):
        """Setting the shape re-draws phi and psi (fading_generators.py:367-386)."""
        self._set_shape(new_shape)
        self._set_phi_and_psi_according_to_shape()

    L = property(lambda self: self._L)
    Ts = property(lambda self: self._Ts)
    Fd = property(lambda self: self._Fd)

    def _set_phi_and_psi_according_to_shape(self):
        """phi is drawn before psi (fading_generators.py:403-425)."""
        dims = [self.L] + (list(self.shape) if self.shape is not None else []) + [1]
        self._phi_l = 2 * np.pi * self.RS.rand(*dims)
        self._psi_l = 2 * np.pi * self.RS.rand(*dims)

    def _advance_clock(self, num_samples):
        """_generate_time_samples (fading_generators.py:427-475) without materialising the vector:
        np.arange(t0, N Ts + t0, Ts 1.0000000001) has values t0 + i*delta with
        delta = (t0 + step) - t0; the new clock is its last value + Ts."""
        t0 = self._current_time
        step = self._Ts * 1.0000000001
        stop = num_samples * self._Ts + t0
        length = int(math.ceil((stop - t0) / step))
        delta = (t0 + step) - t0
        self._current_time = (t0 + (length - 1) * delta) + self._Ts
        return t0, length

    def generate_more_samples(self, num_samples=None):
        """fading_generators.py:495-523."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        n = 1 if num_samples is None else int(num_samples)
        t0, length = self._advance_clock(n)
        P = int(np.prod(self._phi_l.shape[1:]))
        phi = torch.from_numpy(np.ascontiguousarray(self._phi_l.reshape(self._L, P))).cuda()
        psi = torch.from_numpy(np.ascontiguousarray(self._psi_l.reshape(self._L, P))).cuda()
        h = torch.empty((P, length), dtype=torch.complex128, device='cuda')
        _lib.check(lib.b200phy_jakes(_lib.F64, _lib.ptr(phi), _lib.ptr(psi), self._L, P, length,
                                     float(self._Fd), float(self._Ts), float(t0), _lib.ptr(h),
                                     _lib.cur_stream()))
        if self.shape is None:
            h = h.reshape(length)
        else:
            h = h.reshape(tuple(self.shape) + (length,))
        self._store(h)

    def skip_samples_for_next_generation(self, num_samples):
        """fading_generators.py:525-540."""
        self._current_time += num_samples * self._Ts

    def get_similar_fading_generator(self):
        return JakesSampleGenerator(self._Fd, self._Ts, self._L, self._shape)
