"""Fused link ops: one call simulates a batch of independent realizations / frames on the GPU and
returns the four error counters ``[symbol_errors, bit_errors, num_symbols, num_bits]``.

These wrap the ``b200phy_link_*`` entry points of the C ABI.  Two modes:
  * fused mode (``draws=None``): data, channel and noise come from the in-kernel Philox stream,
    a pure function of ``(seed, first_unit + i)`` — invariant to batch size and GPU sharding;
  * stream mode (``draws=(...)``): the draws are CUDA tensors in the documented layouts (what the
    reference's stage API implies; used for parity against the oracle on identical numbers).
``draw_*`` return the fused-mode draws as tensors in exactly those layouts.
"""
import ctypes as C
import os

import numpy as np

from . import SEED_DEFAULT, _lib


MIMO_SCHEMES = {'svd': 1, 'gmd': 2, 'mrt': 3}


def _modem(modulator, dtype):
    return modulator._native(dtype)


def _counters(torch, counters):
    if counters is None:
        return torch.zeros(4, dtype=torch.int64, device='cuda'), True
    return counters, False


def _finish(cnt, own, outs):
    res = cnt.cpu().numpy() if own else cnt
    outs = [o for o in outs if o is not None]
    return (res, *outs) if outs else res


class OfdmTdlLink:
    """Parameters of one OFDM-over-Jakes/TDL link (SISO one-tap or Blast ZF/MMSE), mirroring how
    notebooks/TDL_and_OFDM.ipynb cell 32 builds its objects per frame."""

    def __init__(self, modulator, fft_size, cp_size, num_used_subcarriers=None, *, num_ofdm_symbols=1,
                 Nr=1, Nt=1, tap_powers_linear, tap_delays, Fd=10.0, Ts=None, L=20, t0=None,
                 noise_var=0.01, filter_noise_var=None, dtype='f32', jakes_mode='auto',
                 seed=SEED_DEFAULT, use_pair_kernel=True, use_tensor_cores=False):
        self.modulator = modulator
        self.dtype = _lib.parse_dtype(dtype)
        used = fft_size if num_used_subcarriers is None else num_used_subcarriers
        Ts = 1.0 / (15e3 * fft_size) if Ts is None else Ts
        tap_delays = np.asarray(tap_delays, dtype=np.int64)
        tap_powers_linear = np.asarray(tap_powers_linear, dtype=np.float64)
        if tap_delays.size > _lib.MAX_TAPS:
            raise NotImplementedError("at most %d taps" % _lib.MAX_TAPS)
        p = _lib.OfdmTdlParams()
        p.struct_size = C.sizeof(_lib.OfdmTdlParams)
        p.dtype = self.dtype
        p.fft, p.cp, p.used, p.n_sym = fft_size, cp_size, used, num_ofdm_symbols
        p.Nr, p.Nt, p.n_taps, p.L = Nr, Nt, tap_delays.size, L
        p.jakes_mode = {'auto': _lib.JAKES_AUTO, 'recurrence': _lib.JAKES_RECURRENCE,
                        'poly': _lib.JAKES_POLY}[jakes_mode]
        for i in range(tap_delays.size):
            p.delays[i] = int(tap_delays[i])
            p.tap_powers[i] = float(tap_powers_linear[i])
        p.Fd, p.Ts = float(Fd), float(Ts)
        p.t0 = float(Ts if t0 is None else t0)     # the generator's constructor emits one sample
        p.noise_var = float(noise_var)
        p.filter_noise_var = float(noise_var if filter_noise_var is None else filter_noise_var)
        p.seed = seed
        # bit 0: keep to the generic (non-FFMA2-pair) kernel; bit 1: H_k of the 2x2 / fft-1024 link on the tensor cores
        # (tcgen05, 3xTF32) instead of the CUDA cores — parity-green but measured 3 % slower, hence opt-in
        p.reserved = (0 if use_pair_kernel else 1) | (2 if use_tensor_cores or os.environ.get('B200PHY_TC') else 0)
        self.params = p
        self.mem = int(tap_delays[-1])
        self.N = num_ofdm_symbols * (fft_size + cp_size)
        self.n_data = Nt * num_ofdm_symbols * used
        self.P = L * tap_delays.size * Nr * Nt
        self.Nr, self.Nt = Nr, Nt

    def with_dtype(self, dtype):
        """The same link in another arithmetic ('f32' / 'f64'): same parameters, seed and draws layout."""
        import copy
        other = copy.copy(self)
        other.params = _lib.OfdmTdlParams.from_buffer_copy(self.params)
        other.dtype = _lib.parse_dtype(dtype)
        other.params.dtype = other.dtype
        return other

    def set_noise_var(self, noise_var, filter_noise_var=None):
        self.params.noise_var = float(noise_var)
        self.params.filter_noise_var = float(noise_var if filter_noise_var is None else filter_noise_var)

    def bytes_per_frame(self):
        """Algorithmic HBM bytes of stream mode: idx + noise + phases in, idx out."""
        rs = 4 if self.dtype == _lib.F32 else 8
        return self.n_data + self.Nr * (self.N + self.mem) * 2 * rs + 2 * self.P * rs + self.n_data

    # ---- device calls ------------------------------------------------------------------------
    def draw(self, first_unit, n_units):
        """(idx u8[n, n_data], phi[n, P], psi[n, P], noise[n, Nr, N+mem]) of fused mode."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        idx = torch.empty((n_units, self.n_data), dtype=torch.uint8, device='cuda')
        phi = torch.empty((n_units, self.P), dtype=_lib.real_dtype(self.dtype), device='cuda')
        psi = torch.empty_like(phi)
        noise = torch.empty((n_units, self.Nr, self.N + self.mem), dtype=_lib.cplx_dtype(self.dtype),
                            device='cuda')
        bits = int(round(np.log2(self.modulator.M)))
        _lib.check(lib.b200phy_draw_ofdm_tdl(C.byref(self.params), bits, first_unit, n_units,
                                             _lib.ptr(idx), _lib.ptr(phi), _lib.ptr(psi),
                                             _lib.ptr(noise), _lib.cur_stream()))
        return idx, phi, psi, noise

    def run(self, n_units, first_unit=0, draws=None, counters=None, want_idx=False, want_eq=False,
            want_rx=False):
        """Simulate frames [first_unit, first_unit + n_units).  Returns counters (NumPy int64[4], or
        the device tensor passed in — then nothing synchronises) [, idx_hat][, equalised symbols]
        [, demodulated rx samples before detection: complex[n, Nr, n_sym*used]]."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        modem, keep = _modem(self.modulator, self.dtype)
        cnt, own = _counters(torch, counters)
        idx = phi = psi = noise = None
        if draws is not None:
            idx, phi, psi, noise = draws
        hat = torch.empty((n_units, self.n_data), dtype=torch.uint8, device='cuda') if want_idx else None
        eq = torch.empty((n_units, self.n_data), dtype=_lib.cplx_dtype(self.dtype), device='cuda') \
            if want_eq else None
        rx = torch.empty((n_units, self.Nr, self.n_data // self.Nt), dtype=_lib.cplx_dtype(self.dtype),
                         device='cuda') if want_rx else None
        _lib.check(lib.b200phy_link_ofdm_tdl(C.byref(self.params), modem, first_unit, n_units,
                                             _lib.ptr(idx), _lib.ptr(phi), _lib.ptr(psi),
                                             _lib.ptr(noise), _lib.ptr(hat), _lib.ptr(eq), _lib.ptr(rx),
                                             _lib.ptr(cnt), _lib.cur_stream()))
        return _finish(cnt, own, [hat, eq, rx])

    def run_host(self, n_units, first_unit=0, draws=None, want_idx=False):
        """Same through the host-buffer C entry point (``b200phy_link_ofdm_tdl_host``): draws are
        host tensors (pinned for asynchronous copies); H2D/D2H happen inside the call."""
        lib = _lib.load()
        import torch
        table = np.ascontiguousarray(np.asarray(self.modulator.symbols, dtype=np.complex128))
        tp = table.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))
        cnt = np.zeros(4, dtype=np.int64)
        idx = phi = psi = noise = None
        if draws is not None:
            idx, phi, psi, noise = draws
        # pinned straight from torch's caching host allocator (no pageable staging copy)
        hat = torch.empty((n_units, self.n_data), dtype=torch.uint8, pin_memory=True) if want_idx else None
        _lib.check(lib.b200phy_link_ofdm_tdl_host(
            C.byref(self.params), self.modulator._kind, self.modulator.M, tp, first_unit, n_units,
            _lib.ptr(idx), _lib.ptr(phi), _lib.ptr(psi), _lib.ptr(noise), _lib.ptr(hat),
            cnt.ctypes.data_as(C.POINTER(C.c_int64))))
        return (cnt, hat) if want_idx else cnt


# ---- flat links ----------------------------------------------------------------------------------
def link_siso_flat(modulator, noise_var, n_units, *, rayleigh=True, seed=SEED_DEFAULT, first_unit=0,
                   dtype='f32', draws=None, counters=None, want_idx=False, want_samples=False):
    """One symbol per realization over flat Rayleigh (or AWGN) with perfect-CSI equalisation."""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    dt = _lib.parse_dtype(dtype)
    modem, keep = _modem(modulator, dt)
    cnt, own = _counters(torch, counters)
    idx = h = noise = None
    if draws is not None:
        idx, h, noise = draws
    hat = torch.empty(n_units, dtype=torch.uint8, device='cuda') if want_idx else None
    dec = torch.empty(n_units, dtype=_lib.cplx_dtype(dt), device='cuda') if want_samples else None
    _lib.check(lib.b200phy_link_siso_flat(dt, modem, int(bool(rayleigh)), float(noise_var), seed,
                                          first_unit, n_units, _lib.ptr(idx), _lib.ptr(h),
                                          _lib.ptr(noise), _lib.ptr(hat), _lib.ptr(dec), _lib.ptr(cnt),
                                          _lib.cur_stream()))
    return _finish(cnt, own, [hat, dec])


def link_siso_flat_host(modulator, noise_var, n_units, *, rayleigh=True, seed=SEED_DEFAULT,
                        first_unit=0, dtype='f32', draws=None, want_idx=False):
    lib = _lib.load()
    import torch
    dt = _lib.parse_dtype(dtype)
    table = np.ascontiguousarray(np.asarray(modulator.symbols, dtype=np.complex128))
    tp = table.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))
    cnt = np.zeros(4, dtype=np.int64)
    idx = h = noise = None
    if draws is not None:
        idx, h, noise = draws
    hat = torch.empty(n_units, dtype=torch.uint8, pin_memory=True) if want_idx else None
    _lib.check(lib.b200phy_link_siso_flat_host(dt, modulator._kind, modulator.M, tp, int(bool(rayleigh)),
                                               float(noise_var), seed, first_unit, n_units,
                                               _lib.ptr(idx), _lib.ptr(h), _lib.ptr(noise),
                                               _lib.ptr(hat), cnt.ctypes.data_as(C.POINTER(C.c_int64))))
    return (cnt, hat) if want_idx else cnt


def _host_table(modulator):
    table = np.ascontiguousarray(np.asarray(modulator.symbols, dtype=np.complex128))
    return table, table.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))


def _host_out(n_units, per_unit, want_idx):
    if not want_idx:
        return None
    import torch
    return torch.empty((n_units, per_unit), dtype=torch.uint8, pin_memory=True)


def link_alamouti_host(modulator, noise_var, n_units, *, Nr=2, num_symbols=2, seed=SEED_DEFAULT,
                       first_unit=0, dtype='f32', draws=None, want_idx=False):
    """`link_alamouti` through the host-buffer C entry point (``b200phy_link_alamouti_host``): draws
    are host tensors (or None = Monte Carlo mode: parameters in, 32 bytes of counters out)."""
    lib = _lib.load()
    dt = _lib.parse_dtype(dtype)
    table, tp = _host_table(modulator)
    cnt = np.zeros(4, dtype=np.int64)
    idx, H, noise = draws if draws is not None else (None, None, None)
    hat = _host_out(n_units, num_symbols, want_idx)
    _lib.check(lib.b200phy_link_alamouti_host(dt, modulator._kind, modulator.M, tp, Nr, num_symbols,
                                              float(noise_var), seed, first_unit, n_units, _lib.ptr(idx),
                                              _lib.ptr(H), _lib.ptr(noise), _lib.ptr(hat),
                                              cnt.ctypes.data_as(C.POINTER(C.c_int64))))
    return (cnt, hat) if want_idx else cnt


def link_blast_host(modulator, noise_var, n_units, *, Nr=2, Nt=2, num_symbols=1, filter_noise_var=0.0,
                    seed=SEED_DEFAULT, first_unit=0, dtype='f32', draws=None, want_idx=False):
    """`link_blast` through ``b200phy_link_blast_host`` (host draws or Monte Carlo mode)."""
    lib = _lib.load()
    dt = _lib.parse_dtype(dtype)
    table, tp = _host_table(modulator)
    cnt = np.zeros(4, dtype=np.int64)
    idx, H, noise = draws if draws is not None else (None, None, None)
    hat = _host_out(n_units, num_symbols * Nt, want_idx)
    _lib.check(lib.b200phy_link_blast_host(dt, modulator._kind, modulator.M, tp, Nr, Nt, num_symbols,
                                           float(noise_var), float(filter_noise_var), seed, first_unit, n_units,
                                           _lib.ptr(idx), _lib.ptr(H), _lib.ptr(noise), _lib.ptr(hat),
                                           cnt.ctypes.data_as(C.POINTER(C.c_int64))))
    return (cnt, hat) if want_idx else cnt


def link_precoded_host(modulator, noise_var, n_units, *, scheme, Nr, Nt, num_symbols=1, filter_noise_var=0.0,
                       seed=SEED_DEFAULT, first_unit=0, dtype='f32', draws=None, want_idx=False):
    """`link_precoded` through ``b200phy_link_precoded_host`` (host draws or Monte Carlo mode)."""
    lib = _lib.load()
    dt = _lib.parse_dtype(dtype)
    if scheme not in MIMO_SCHEMES:
        raise ValueError("scheme must be one of %s" % sorted(MIMO_SCHEMES))
    table, tp = _host_table(modulator)
    cnt = np.zeros(4, dtype=np.int64)
    idx, H, noise = draws if draws is not None else (None, None, None)
    layers = 1 if scheme == 'mrt' else Nt
    hat = _host_out(n_units, num_symbols * layers, want_idx)
    _lib.check(lib.b200phy_link_precoded_host(dt, modulator._kind, modulator.M, tp, MIMO_SCHEMES[scheme], Nr, Nt,
                                              num_symbols, float(noise_var), float(filter_noise_var), seed,
                                              first_unit, n_units, _lib.ptr(idx), _lib.ptr(H), _lib.ptr(noise),
                                              _lib.ptr(hat), cnt.ctypes.data_as(C.POINTER(C.c_int64))))
    return (cnt, hat) if want_idx else cnt


def draw_siso_flat(modulator, n_units, *, rayleigh=True, seed=SEED_DEFAULT, first_unit=0, dtype='f32'):
    lib = _lib.load()
    torch = _lib.torch_cuda()
    dt = _lib.parse_dtype(dtype)
    idx = torch.empty(n_units, dtype=torch.uint8, device='cuda')
    h = torch.empty(n_units, dtype=_lib.cplx_dtype(dt), device='cuda') if rayleigh else None
    noise = torch.empty(n_units, dtype=_lib.cplx_dtype(dt), device='cuda')
    bits = int(round(np.log2(modulator.M)))
    _lib.check(lib.b200phy_draw_siso_flat(dt, bits, seed, first_unit, n_units, _lib.ptr(idx),
                                          _lib.ptr(h), _lib.ptr(noise), _lib.cur_stream()))
    return idx, h, noise


def link_alamouti(modulator, noise_var, n_units, *, Nr=2, num_symbols=2, seed=SEED_DEFAULT,
                  first_unit=0, dtype='f32', draws=None, counters=None, want_idx=False,
                  want_samples=False):
    """Alamouti over flat Rayleigh: per realization H[Nr, 2], `num_symbols` symbols."""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    dt = _lib.parse_dtype(dtype)
    modem, keep = _modem(modulator, dt)
    cnt, own = _counters(torch, counters)
    idx = H = noise = None
    if draws is not None:
        idx, H, noise = draws
    S = num_symbols
    hat = torch.empty((n_units, S), dtype=torch.uint8, device='cuda') if want_idx else None
    dec = torch.empty((n_units, S), dtype=_lib.cplx_dtype(dt), device='cuda') if want_samples else None
    _lib.check(lib.b200phy_link_alamouti(dt, modem, Nr, S, float(noise_var), seed, first_unit, n_units,
                                         _lib.ptr(idx), _lib.ptr(H), _lib.ptr(noise), _lib.ptr(hat),
                                         _lib.ptr(dec), _lib.ptr(cnt), _lib.cur_stream()))
    return _finish(cnt, own, [hat, dec])


def link_blast(modulator, noise_var, n_units, *, Nr=2, Nt=2, num_symbols=1, filter_noise_var=0.0,
               seed=SEED_DEFAULT, first_unit=0, dtype='f32', draws=None, counters=None,
               want_idx=False, want_samples=False):
    """Blast (V-BLAST spatial multiplexing) with a ZF (filter_noise_var=0) or MMSE receive filter;
    `num_symbols` symbol vectors (Nt symbols each) per realization."""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    dt = _lib.parse_dtype(dtype)
    modem, keep = _modem(modulator, dt)
    cnt, own = _counters(torch, counters)
    idx = H = noise = None
    if draws is not None:
        idx, H, noise = draws
    S = num_symbols
    hat = torch.empty((n_units, S * Nt), dtype=torch.uint8, device='cuda') if want_idx else None
    dec = torch.empty((n_units, S * Nt), dtype=_lib.cplx_dtype(dt), device='cuda') if want_samples else None
    _lib.check(lib.b200phy_link_blast(dt, modem, Nr, Nt, S, float(noise_var), float(filter_noise_var),
                                      seed, first_unit, n_units, _lib.ptr(idx), _lib.ptr(H),
                                      _lib.ptr(noise), _lib.ptr(hat), _lib.ptr(dec), _lib.ptr(cnt),
                                      _lib.cur_stream()))
    return _finish(cnt, own, [hat, dec])


def link_precoded(modulator, noise_var, n_units, *, scheme, Nr, Nt, num_symbols=1, filter_noise_var=0.0,
                  seed=SEED_DEFAULT, first_unit=0, dtype='f32', draws=None, counters=None,
                  want_idx=False, want_samples=False):
    """SVDMimo / GMDMimo (square Nr == Nt, Nt layers) or MRT (Nr == 1, one layer) over flat Rayleigh:
    one channel realization per unit, decomposed on the GPU, `num_symbols` symbol vectors through
    precoder -> channel -> receive filter.  Symbol p = l * num_symbols + s is layer l at time s."""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    dt = _lib.parse_dtype(dtype)
    if scheme not in MIMO_SCHEMES:
        raise ValueError("scheme must be one of %s" % sorted(MIMO_SCHEMES))
    modem, keep = _modem(modulator, dt)
    cnt, own = _counters(torch, counters)
    idx = H = noise = None
    if draws is not None:
        idx, H, noise = draws
    S = num_symbols
    layers = 1 if scheme == 'mrt' else Nt
    hat = torch.empty((n_units, S * layers), dtype=torch.uint8, device='cuda') if want_idx else None
    dec = torch.empty((n_units, S * layers), dtype=_lib.cplx_dtype(dt), device='cuda') if want_samples else None
    _lib.check(lib.b200phy_link_precoded(dt, modem, MIMO_SCHEMES[scheme], Nr, Nt, S, float(noise_var),
                                         float(filter_noise_var), seed, first_unit, n_units, _lib.ptr(idx),
                                         _lib.ptr(H), _lib.ptr(noise), _lib.ptr(hat), _lib.ptr(dec),
                                         _lib.ptr(cnt), _lib.cur_stream()))
    return _finish(cnt, own, [hat, dec])


def draw_flat_mimo(modulator, n_units, *, Nr, Nt, num_symbols, n_data, seed=SEED_DEFAULT, first_unit=0,
                   dtype='f32'):
    """(idx u8[n, n_data], H[n, Nr, Nt], noise[n, Nr, num_symbols]) of fused mode."""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    dt = _lib.parse_dtype(dtype)
    idx = torch.empty((n_units, n_data), dtype=torch.uint8, device='cuda')
    H = torch.empty((n_units, Nr, Nt), dtype=_lib.cplx_dtype(dt), device='cuda')
    noise = torch.empty((n_units, Nr, num_symbols), dtype=_lib.cplx_dtype(dt), device='cuda')
    bits = int(round(np.log2(modulator.M)))
    _lib.check(lib.b200phy_draw_flat_mimo(dt, bits, Nr, Nt, num_symbols, n_data, seed, first_unit,
                                          n_units, _lib.ptr(idx), _lib.ptr(H), _lib.ptr(noise),
                                          _lib.cur_stream()))
    return idx, H, noise
