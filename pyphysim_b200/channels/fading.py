"""Tapped-delay-line channel with the API of pyphysim.channels.fading; per-sample work on the GPU.

Host side (one-off): channel profiles and their discretisation.
Device side: fading-sample generation (``b200phy_jakes`` / Philox Rayleigh), the time-varying sparse
FIR (``b200phy_tdl_apply``) and the frequency response (``b200phy_tdl_freq_response``).
"""
import math

import numpy as np

from .. import _device as D
from .. import _lib
from ..util.conversion import dB2Linear, linear2dB
from . import fading_generators as fg

__all__ = ['TdlChannelProfile', 'TdlImpulseResponse', 'TdlChannel', 'TdlMimoChannel',
           'COST259_TUx', 'COST259_RAx', 'COST259_HTx']


class TdlChannelProfile:
    """Powers and delays of the taps of a TDL profile (reference: channels/fading.py:28-316)."""

    def __init__(self, tap_powers_dB=None, tap_delays=None, name='custom'):
        self._name = name
        if tap_powers_dB is None and tap_delays is None:
            tap_powers_dB = np.zeros(1)
            tap_delays = np.zeros(1)
        self._tap_powers_dB = np.array(tap_powers_dB, dtype=float)
        self._tap_powers_dB.flags['WRITEABLE'] = False
        self._tap_powers_linear = dB2Linear(self._tap_powers_dB)
        self._tap_powers_linear.flags['WRITEABLE'] = False
        self._tap_delays = np.array(tap_delays)
        self._tap_delays.flags['WRITEABLE'] = False
        self._num_taps = self._tap_delays.size
        p, d = self._tap_powers_linear, self._tap_delays
        self._mean_excess_delay = float(np.sum(p * d) / np.sum(p))
        self._rms_delay_spread = math.sqrt(float(np.sum(p * d ** 2) / np.sum(p))
                                           - self._mean_excess_delay ** 2)
        self._Ts = None

    mean_excess_delay = property(lambda self: self._mean_excess_delay)
    rms_delay_spread = property(lambda self: self._rms_delay_spread)
    name = property(lambda self: self._name)
    tap_powers_dB = property(lambda self: self._tap_powers_dB)
    tap_powers_linear = property(lambda self: self._tap_powers_linear)
    tap_delays = property(lambda self: self._tap_delays)
    num_taps = property(lambda self: self._num_taps)
    Ts = property(lambda self: self._Ts)

    @property
    def num_taps_with_padding(self):
        if self._Ts is None:
            raise RuntimeError('TdlChannelProfile is not discretized')
        return int(self._tap_delays[-1] + 1)

    @property
    def is_discretized(self):
        return self._Ts is not None

    def get_discretize_profile(self, Ts):
        """fading.py:236-304: round delays to samples, sum colliding taps (linear), normalise."""
        if self.is_discretized:
            raise RuntimeError("Trying to discretize a TdlChannelProfile "
                               "object that is already discretized.")
        powers, delays = self._calc_discretized_tap_powers_and_delays(Ts)
        out = TdlChannelProfile(powers, delays, "{0} (discretized)".format(self.name))
        out._Ts = Ts
        return out

    def _calc_discretized_tap_powers_and_delays(self, Ts):
        slots = np.round(self._tap_delays / Ts).astype(int).flatten()
        delay_indexes, inverse = np.unique(slots, return_inverse=True)
        acc = np.zeros(delay_indexes.size)
        np.add.at(acc, inverse, self._tap_powers_linear)
        acc /= np.sum(acc)
        return linear2dB(acc), delay_indexes

    def __repr__(self):
        return "<TdlChannelProfile: '{0}' ({1} taps)>".format(self.name, self.num_taps)


# 3GPP TR 25.943 (COST 259) profiles, values as tabulated in channels/fading.py:327-353
COST259_TUx = TdlChannelProfile(
    np.array([-5.7, -7.6, -10.1, -10.2, -10.2, -11.5, -13.4, -16.3, -16.9, -17.1,
              -17.4, -19, -19, -19.8, -21.5, -21.6, -22.1, -22.6, -23.5, -24.3]),
    np.array([0, 217, 512, 514, 517, 674, 882, 1230, 1287, 1311, 1349, 1533, 1535,
              1622, 1818, 1836, 1884, 1943, 2048, 2140]) * 1e-9, 'COST259_TU')
COST259_RAx = TdlChannelProfile(
    np.array([-5.2, -6.4, -8.4, -9.3, -10.0, -13.1, -15.3, -18.5, -20.4, -22.4]),
    np.array([0., 42., 101., 129., 149., 245., 312., 410., 469., 528]) * 1e-9, 'COST259_RA')
COST259_HTx = TdlChannelProfile(
    np.array([-3.6, -8.9, -10.2, -11.5, -11.8, -12.7, -13.0, -16.2, -17.3, -17.7,
              -17.6, -22.7, -24.1, -25.8, -25.8, -26.2, -29.0, -29.9, -30.0, -30.7]),
    np.array([0., 356., 441., 528., 546., 609., 625., 842., 916., 941., 15000.,
              16172., 16492., 16876., 16882., 16978., 17615., 17827., 17849., 18016.]) * 1e-9,
    'COST259_HT')


def _scale_rows(t, scales):
    """t[row, ...] *= scales[row] on the device (b200phy_scale_rows)."""
    import ctypes as C
    lib = _lib.load()
    sc = np.ascontiguousarray(scales, dtype=np.float64)
    rows = sc.size
    _lib.check(lib.b200phy_scale_rows(_lib.F64, _lib.ptr(t), rows, t.numel() // rows,
                                      sc.ctypes.data_as(C.POINTER(C.c_double)), _lib.cur_stream()))
    return t


class TdlImpulseResponse:
    """Impulse response of a TDL channel over time (reference: fading.py:356-698).  The sparse tap
    values live on the device; NumPy views are produced on demand."""

    def __init__(self, tap_values, channel_profile):
        assert isinstance(channel_profile, TdlChannelProfile)
        if channel_profile.Ts is None:
            raise RuntimeError('Channel profile must be discretized')
        self._channel_profile = channel_profile
        if D.is_torch(tap_values):
            self._sparse_dev, self._sparse_np = tap_values, None
        else:
            self._sparse_np = np.asarray(tap_values, dtype=complex)
            self._sparse_dev = None
        self._tap_values_dense = None

    def _dev(self):
        if self._sparse_dev is None:
            self._sparse_dev, _ = D.to_device(self._sparse_np, np.complex128)
        return self._sparse_dev

    @property
    def tap_values_sparse(self):
        if self._sparse_np is None:
            self._sparse_np = self._sparse_dev.cpu().numpy()
        return self._sparse_np

    @property
    def tap_indexes_sparse(self):
        return self._channel_profile.tap_delays

    @property
    def Ts(self):
        return self._channel_profile.Ts

    @property
    def tap_delays_sparse(self):
        return self.tap_indexes_sparse * self.Ts

    @property
    def tap_values(self):
        """Dense taps, zero-filled, read-only (fading.py:482-511)."""
        if self._tap_values_dense is None:
            sp = self.tap_values_sparse
            dense = np.zeros((int(self.tap_indexes_sparse[-1]) + 1,) + sp.shape[1:], dtype=complex)
            dense[self.tap_indexes_sparse] = sp
            dense.flags['WRITEABLE'] = False
            self._tap_values_dense = dense
        return self._tap_values_dense

    @property
    def num_samples(self):
        shape = self._sparse_dev.shape if self._sparse_dev is not None else self._sparse_np.shape
        return int(shape[-1])

    @property
    def channel_profile(self):
        return self._channel_profile

    def get_freq_response(self, fft_size):
        """fading.py:513-536: FFT over the delay axis for every time sample -> [fft, ..., N]."""
        import ctypes as C
        lib = _lib.load()
        torch = _lib.torch_cuda()
        taps = self._dev()
        delays = np.ascontiguousarray(self.tap_indexes_sparse, dtype=np.int32)
        n_taps, N = taps.shape[0], taps.shape[-1]
        A = taps.numel() // (n_taps * N)
        out = torch.empty((fft_size,) + tuple(taps.shape[1:]), dtype=torch.complex128, device='cuda')
        _lib.check(lib.b200phy_tdl_freq_response(_lib.F64, _lib.ptr(taps),
                                                 delays.ctypes.data_as(C.POINTER(C.c_int32)), n_taps, A, N,
                                                 int(fft_size), _lib.ptr(out), _lib.cur_stream()))
        return out.cpu().numpy()

    def __mul__(self, value):
        """fading.py:538-559."""
        t = self._dev().clone()
        _scale_rows(t, [float(value)])
        return TdlImpulseResponse(t, self._channel_profile)

    def __rmul__(self, value):
        return self * value

    @staticmethod
    def concatenate_samples(impulse_responses):
        """fading.py:655-698."""
        num = len(impulse_responses)
        if num < 2:
            if num == 1:
                return impulse_responses[0]
            raise ValueError("impulse_responses must contain at least two TdlImpulseResponse objects.")
        if impulse_responses[0].channel_profile is not impulse_responses[1].channel_profile:
            raise ValueError("TdlImpulseResponse objects must have the same channel profile object")
        import torch
        taps = torch.cat([a._dev() for a in impulse_responses], dim=-1)
        return TdlImpulseResponse(taps, impulse_responses[0].channel_profile)


class TdlChannel:
    """Tapped-delay-line channel (reference: fading.py:701-1287)."""

    def __init__(self, fading_generator, channel_profile=None, tap_powers_dB=None, tap_delays=None,
                 Ts=None):
        if isinstance(fading_generator, fg.JakesSampleGenerator):
            if Ts is None:
                Ts = fading_generator.Ts
            elif Ts != fading_generator.Ts:
                raise RuntimeError("The provided sampling interval Ts is different from "
                                   "the one in the Jakes sample generator.")
        if channel_profile is None:
            channel_profile = TdlChannelProfile(tap_powers_dB, tap_delays)
        else:
            assert isinstance(channel_profile, TdlChannelProfile), \
                'channel_profile must be an obj of the TdlChannelProfile class'
        if not channel_profile.is_discretized:
            if isinstance(fading_generator, fg.RayleighSampleGenerator) and Ts is None:
                Ts = 1.0
            assert Ts is not None
            channel_profile = channel_profile.get_discretize_profile(Ts)
        elif channel_profile.Ts != Ts and Ts is not None:
            raise RuntimeError("Channel profile is already discretized, but it does not "
                               "agree with the discretized parameter Ts")
        self._channel_profile = channel_profile
        self._fading_generator = fading_generator
        self._set_fading_generator_shape(fading_generator.shape)
        self._last_impulse_response = None
        self._switched_direction = False

    @property
    def switched_direction(self):
        return self._switched_direction

    @switched_direction.setter
    def switched_direction(self, value):
        if not isinstance(value, bool):
            raise TypeError("switched_direction must be a boolean value")
        self._switched_direction = value

    def set_num_antennas(self, num_rx_antennas, num_tx_antennas):
        self._set_fading_generator_shape((num_rx_antennas, num_tx_antennas))

    def _set_fading_generator_shape(self, new_shape):
        if new_shape is None:
            self._fading_generator.shape = (self.num_taps,)
        else:
            self._fading_generator.shape = (self.num_taps,) + tuple(new_shape)

    channel_profile = property(lambda self: self._channel_profile)
    num_taps = property(lambda self: self._channel_profile.num_taps)
    num_taps_with_padding = property(lambda self: self._channel_profile.num_taps_with_padding)

    @property
    def num_tx_antennas(self):
        s = self._fading_generator.shape
        return -1 if s is None or len(s) == 1 else s[2]

    @property
    def num_rx_antennas(self):
        s = self._fading_generator.shape
        return -1 if s is None or len(s) == 1 else s[1]

    def generate_impulse_response(self, num_samples=1):
        """fading.py:908-959: fading samples * sqrt(tap power), kept on the device."""
        self._fading_generator.generate_more_samples(num_samples)
        samples = self._fading_generator._device_samples().clone()
        if samples.dim() == len(self._fading_generator.shape):      # num_samples axis missing
            samples = samples.unsqueeze(-1)
        _scale_rows(samples, np.sqrt(self._channel_profile.tap_powers_linear))
        self._last_impulse_response = TdlImpulseResponse(samples, self._channel_profile)

    def get_last_impulse_response(self):
        if self._last_impulse_response is None:
            raise RuntimeError("No impulse response was generated yet")
        return self._last_impulse_response

    def _prepare_transmit_signal_shape(self, signal):
        """fading.py:1009-1044: 1-D input is accepted for a single transmit antenna."""
        shape = self._fading_generator.shape
        if len(shape) == 1:
            return signal
        _, num_rx, num_tx = shape
        single = num_rx if self.switched_direction else num_tx
        if single == 1 and signal.dim() == 1:
            signal = signal.reshape(1, -1)
        return signal

    def corrupt_data(self, signal):
        """fading.py:1046-1124: time-varying sparse FIR (b200phy_tdl_apply)."""
        import ctypes as C
        lib = _lib.load()
        torch = _lib.torch_cuda()
        x, was_np = D.to_device(signal, np.complex128)
        num_symbols = x.shape[-1]
        x = self._prepare_transmit_signal_shape(x)
        self.generate_impulse_response(num_symbols)
        taps = self._last_impulse_response._dev()
        delays = np.ascontiguousarray(self._channel_profile.tap_delays, dtype=np.int32)
        ones = np.ones(delays.size)
        mem = self.num_taps_with_padding - 1
        shape = self._fading_generator.shape
        if len(shape) == 1:
            Nr = Nt = 1
            out_shape = (num_symbols + mem,)
        elif len(shape) == 3:
            _, num_rx, num_tx = shape
            if self.switched_direction:
                taps = taps.permute(0, 2, 1, 3).contiguous()    # roles of the antennas swap
                Nr, Nt = num_tx, num_rx
            else:
                Nr, Nt = num_rx, num_tx
            out_shape = (Nr, num_symbols + mem)
            if x.dim() != 2 or x.shape[0] != Nt:
                raise ValueError("signal must have shape (%d, num_samples)" % Nt)
        else:  # pragma: no cover
            raise RuntimeError("Shape of the fading generator of the TdlChannel class must "
                               "have either 1 (SISO) or 3 (MIMO) dimensions")
        y = torch.empty(out_shape, dtype=torch.complex128, device='cuda')
        _lib.check(lib.b200phy_tdl_apply(_lib.F64, _lib.ptr(x.contiguous()), _lib.ptr(taps),
                                         ones.ctypes.data_as(C.POINTER(C.c_double)),
                                         delays.ctypes.data_as(C.POINTER(C.c_int32)), delays.size, Nr, Nt,
                                         num_symbols, _lib.ptr(y), _lib.cur_stream()))
        return D.from_device(y, was_np)

    def _generate_block_impulse_responses(self, num_blocks, fft_size):
        """One impulse-response sample per block, the generator advancing `fft_size` samples per
        block (generate 1 + skip fft_size-1, fading.py:1203-1276), in ONE device call."""
        gen = self._fading_generator
        if isinstance(gen, fg.JakesSampleGenerator):
            lib = _lib.load()
            torch = _lib.torch_cuda()
            t0 = gen._current_time
            P = int(np.prod(gen._phi_l.shape[1:]))
            phi = torch.from_numpy(np.ascontiguousarray(gen._phi_l.reshape(gen.L, P))).cuda()
            psi = torch.from_numpy(np.ascontiguousarray(gen._psi_l.reshape(gen.L, P))).cuda()
            h = torch.empty((P, num_blocks), dtype=torch.complex128, device='cuda')
            # block times t0 + b*fft*Ts: the kernel's grid is t0 + n*Ts_eff*1.0000000001
            _lib.check(lib.b200phy_jakes(_lib.F64, _lib.ptr(phi), _lib.ptr(psi), gen.L, P, num_blocks,
                                         float(gen.Fd), float(gen.Ts) * fft_size / 1.0000000001, float(t0),
                                         _lib.ptr(h), _lib.cur_stream()))
            for _ in range(num_blocks):                 # replay the reference's clock arithmetic
                gen._advance_clock(1)
                gen.skip_samples_for_next_generation(fft_size - 1)
            samples = h.reshape(tuple(gen.shape) + (num_blocks,))
            gen._store(samples[..., -1:].clone())
        else:
            gen.generate_more_samples(num_blocks)
            samples = gen._device_samples()
        samples = samples.clone()
        _scale_rows(samples, np.sqrt(self._channel_profile.tap_powers_linear))
        self._last_impulse_response = TdlImpulseResponse(samples, self._channel_profile)
        return samples

    def corrupt_data_in_freq_domain(self, signal, fft_size, carrier_indexes=None):
        """Block-static frequency-domain path (fading.py:1126-1287): every block of `fft_size` (or
        `len(carrier_indexes)`) symbols is multiplied by the frequency response of one impulse response,
        and the fading generator advances `fft_size` samples per block."""
        import ctypes as C
        lib = _lib.load()
        torch = _lib.torch_cuda()
        x, was_np = D.to_device(signal, np.complex128)
        num_symbols = x.shape[-1]
        x = self._prepare_transmit_signal_shape(x)
        if carrier_indexes is None:
            block_size, car = fft_size, None
        elif isinstance(carrier_indexes, slice):
            start, stop, step = carrier_indexes.indices(fft_size)
            block_size = (stop - start) // step         # same expression as the reference (:1170-1172)
            car = np.arange(fft_size)[carrier_indexes][:block_size]
        else:
            car = np.asarray(carrier_indexes)
            block_size = len(car)
        if num_symbols % block_size != 0:
            raise ValueError("The num of elements in `signal` must be a multiple of number of sent "
                             "elements per `fft_size`.")
        num_blocks = num_symbols // block_size
        shape = self._fading_generator.shape
        taps = self._generate_block_impulse_responses(num_blocks, fft_size)
        if len(shape) == 1:
            Nr = Nt = 1
        else:
            _, num_rx, num_tx = shape
            if self.switched_direction:
                taps = taps.permute(0, 2, 1, 3).contiguous()
                Nr, Nt = num_tx, num_rx
            else:
                Nr, Nt = num_rx, num_tx
            if x.dim() != 2 or x.shape[0] != Nt:
                raise ValueError("signal must have shape (%d, num_samples)" % Nt)
        delays = np.ascontiguousarray(self._channel_profile.tap_delays, dtype=np.int32)
        H = torch.empty((fft_size, Nr, Nt, num_blocks), dtype=torch.complex128, device='cuda')
        _lib.check(lib.b200phy_tdl_freq_response(_lib.F64, _lib.ptr(taps),
                                                 delays.ctypes.data_as(C.POINTER(C.c_int32)), delays.size,
                                                 Nr * Nt, num_blocks, int(fft_size), _lib.ptr(H),
                                                 _lib.cur_stream()))
        car_dev = None
        if car is not None:
            car_dev = torch.from_numpy(np.ascontiguousarray(np.mod(car, fft_size), dtype=np.int32)).cuda()
        y = torch.empty((Nr, num_symbols), dtype=torch.complex128, device='cuda')
        _lib.check(lib.b200phy_freq_apply(_lib.F64, _lib.ptr(H), _lib.ptr(x.contiguous()), _lib.ptr(car_dev),
                                          int(fft_size), int(block_size), Nr, Nt, num_blocks, _lib.ptr(y),
                                          _lib.cur_stream()))
        if len(shape) == 1:
            y = y.reshape(num_symbols)
        return D.from_device(y, was_np)


class TdlMimoChannel(TdlChannel):
    """fading.py:1290-1333."""

    def __init__(self, fading_generator, channel_profile=None, tap_powers_dB=None, tap_delays=None,
                 Ts=None):
        if fading_generator.shape is None or len(fading_generator.shape) != 2:
            raise RuntimeError("The provided fading_generator for the TdlMimoChannel class"
                               " must have a shape with two values")
        super().__init__(fading_generator, channel_profile, tap_powers_dB, tap_delays, Ts)
