"""``torch.ops.b200phy.*`` — the fused link kernels as registered PyTorch operators.

SURVEY.md §8(b) asks for the throughput path to be reachable as torch ops whose only cross-boundary types are
tensors and plain scalars.  The product boundary is the C ABI of ``libb200phy.so`` (``include/b200phy.h``: raw
pointers, sizes and a stream, no torch types — what a cgo/JNI/ctypes binding needs); this module is the thin
adapter above it: every op is declared with a schema in the ``b200phy`` namespace (``torch.library``), takes CUDA
tensors, launches on the current stream and accumulates into the ``counters`` tensor in place
(``[symbol_errors, bit_errors, num_symbols, num_bits]``, int64[4]).  Draw tensors are optional: ``None`` = Monte
Carlo mode (in-kernel Philox stream keyed by ``(seed, first_unit + i)``).  The arithmetic follows the table's
dtype (complex64 -> f32, complex128 -> f64).  There is no CPU kernel: calling an op with CPU tensors raises.

    import pyphysim_b200.torch_ops          # registers the ops
    cnt = torch.zeros(4, dtype=torch.int64, device='cuda')
    torch.ops.b200phy.link_siso_flat(table, KIND_QAM, True, 0.01, seed, 0, 1_000_000, None, None, None, cnt)

Reference operations behind the ops: notebook Rayleigh cell 8 (siso_flat), apps/mimo/simulate_mimo.py:68-142
(alamouti / blast / precoded), notebooks/TDL_and_OFDM.ipynb cell 32 + modulators/ofdm.py:394-552 (ofdm_tdl).
"""
import ctypes as C

import torch

from . import _lib

_MASK64 = (1 << 64) - 1

_DEFS = {
    'link_siso_flat':
        '(Tensor table, int kind, bool rayleigh, float noise_var, int seed, int first_unit, int n_units, '
        'Tensor? idx, Tensor? h, Tensor? noise, Tensor(a!) counters, Tensor(b!)? idx_hat=None) -> ()',
    'link_alamouti':
        '(Tensor table, int kind, int Nr, int num_symbols, float noise_var, int seed, int first_unit, int n_units, '
        'Tensor? idx, Tensor? H, Tensor? noise, Tensor(a!) counters, Tensor(b!)? idx_hat=None) -> ()',
    'link_blast':
        '(Tensor table, int kind, int Nr, int Nt, int num_symbols, float noise_var, float filter_noise_var, int seed, '
        'int first_unit, int n_units, Tensor? idx, Tensor? H, Tensor? noise, Tensor(a!) counters, '
        'Tensor(b!)? idx_hat=None) -> ()',
    'link_precoded':
        '(Tensor table, int kind, int scheme, int Nr, int Nt, int num_symbols, float noise_var, float filter_noise_var, '
        'int seed, int first_unit, int n_units, Tensor? idx, Tensor? H, Tensor? noise, Tensor(a!) counters, '
        'Tensor(b!)? idx_hat=None) -> ()',
    'link_ofdm_tdl':
        '(Tensor table, int kind, int fft, int cp, int used, int n_sym, int Nr, int Nt, int[] delays, '
        'float[] tap_powers, float Fd, float Ts, float t0, int L, float noise_var, float filter_noise_var, int seed, '
        'int first_unit, int n_units, Tensor? idx, Tensor? phi, Tensor? psi, Tensor? noise, Tensor(a!) counters, '
        'Tensor(b!)? idx_hat=None) -> ()',
}

_library = torch.library.Library('b200phy', 'DEF')
for _name, _schema in _DEFS.items():
    _library.define(_name + _schema)


def _dtype_of(table):
    if table.dtype == torch.complex64:
        return _lib.F32
    if table.dtype == torch.complex128:
        return _lib.F64
    raise ValueError('table must be complex64 or complex128, got %s' % table.dtype)


def _modem(table, kind):
    if not table.is_cuda:
        raise _lib.B200PhyError('b200phy ops need CUDA tensors (B200, sm_100a); there is no CPU kernel')
    if not table.is_contiguous():
        raise ValueError('table must be contiguous')
    return _lib.Modem(int(kind), int(table.numel()), table.data_ptr())


def _check(counters, *tensors):
    if counters.dtype != torch.int64 or counters.numel() != 4 or not counters.is_cuda:
        raise ValueError('counters must be a CUDA int64[4] tensor')
    for t in tensors:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise ValueError('draw / output tensors must be contiguous CUDA tensors')


def _u64(v):
    return int(v) & _MASK64


def link_siso_flat(table, kind, rayleigh, noise_var, seed, first_unit, n_units, idx, h, noise, counters,
                   idx_hat=None):
    _check(counters, idx, h, noise, idx_hat)
    lib = _lib.load()
    m = _modem(table, kind)
    _lib.check(lib.b200phy_link_siso_flat(_dtype_of(table), C.byref(m), int(bool(rayleigh)), float(noise_var),
                                          _u64(seed), _u64(first_unit), int(n_units), _lib.ptr(idx), _lib.ptr(h),
                                          _lib.ptr(noise), _lib.ptr(idx_hat), None, _lib.ptr(counters),
                                          _lib.cur_stream()))


def link_alamouti(table, kind, Nr, num_symbols, noise_var, seed, first_unit, n_units, idx, H, noise, counters,
                  idx_hat=None):
    _check(counters, idx, H, noise, idx_hat)
    lib = _lib.load()
    m = _modem(table, kind)
    _lib.check(lib.b200phy_link_alamouti(_dtype_of(table), C.byref(m), int(Nr), int(num_symbols), float(noise_var),
                                         _u64(seed), _u64(first_unit), int(n_units), _lib.ptr(idx), _lib.ptr(H),
                                         _lib.ptr(noise), _lib.ptr(idx_hat), None, _lib.ptr(counters),
                                         _lib.cur_stream()))


def link_blast(table, kind, Nr, Nt, num_symbols, noise_var, filter_noise_var, seed, first_unit, n_units, idx, H,
               noise, counters, idx_hat=None):
    _check(counters, idx, H, noise, idx_hat)
    lib = _lib.load()
    m = _modem(table, kind)
    _lib.check(lib.b200phy_link_blast(_dtype_of(table), C.byref(m), int(Nr), int(Nt), int(num_symbols),
                                      float(noise_var), float(filter_noise_var), _u64(seed), _u64(first_unit),
                                      int(n_units), _lib.ptr(idx), _lib.ptr(H), _lib.ptr(noise), _lib.ptr(idx_hat),
                                      None, _lib.ptr(counters), _lib.cur_stream()))


def link_precoded(table, kind, scheme, Nr, Nt, num_symbols, noise_var, filter_noise_var, seed, first_unit, n_units,
                  idx, H, noise, counters, idx_hat=None):
    _check(counters, idx, H, noise, idx_hat)
    lib = _lib.load()
    m = _modem(table, kind)
    _lib.check(lib.b200phy_link_precoded(_dtype_of(table), C.byref(m), int(scheme), int(Nr), int(Nt),
                                         int(num_symbols), float(noise_var), float(filter_noise_var), _u64(seed),
                                         _u64(first_unit), int(n_units), _lib.ptr(idx), _lib.ptr(H), _lib.ptr(noise),
                                         _lib.ptr(idx_hat), None, _lib.ptr(counters), _lib.cur_stream()))


def link_ofdm_tdl(table, kind, fft, cp, used, n_sym, Nr, Nt, delays, tap_powers, Fd, Ts, t0, L, noise_var,
                  filter_noise_var, seed, first_unit, n_units, idx, phi, psi, noise, counters, idx_hat=None):
    _check(counters, idx, phi, psi, noise, idx_hat)
    if len(delays) != len(tap_powers) or not 0 < len(delays) <= _lib.MAX_TAPS:
        raise ValueError('delays / tap_powers: 1..%d taps of equal count' % _lib.MAX_TAPS)
    lib = _lib.load()
    m = _modem(table, kind)
    p = _lib.OfdmTdlParams()
    p.struct_size = C.sizeof(_lib.OfdmTdlParams)
    p.dtype = _dtype_of(table)
    p.fft, p.cp, p.used, p.n_sym = int(fft), int(cp), int(used), int(n_sym)
    p.Nr, p.Nt, p.n_taps, p.L = int(Nr), int(Nt), len(delays), int(L)
    p.jakes_mode = _lib.JAKES_AUTO
    for i, (d, w) in enumerate(zip(delays, tap_powers)):
        p.delays[i] = int(d)
        p.tap_powers[i] = float(w)
    p.Fd, p.Ts, p.t0 = float(Fd), float(Ts), float(t0)
    p.noise_var, p.filter_noise_var = float(noise_var), float(filter_noise_var)
    p.seed = _u64(seed)
    _lib.check(lib.b200phy_link_ofdm_tdl(C.byref(p), C.byref(m), _u64(first_unit), int(n_units), _lib.ptr(idx),
                                         _lib.ptr(phi), _lib.ptr(psi), _lib.ptr(noise), _lib.ptr(idx_hat), None, None,
                                         _lib.ptr(counters), _lib.cur_stream()))


_IMPLS = {'link_siso_flat': link_siso_flat, 'link_alamouti': link_alamouti, 'link_blast': link_blast,
          'link_precoded': link_precoded, 'link_ofdm_tdl': link_ofdm_tdl}
for _name, _fn in _IMPLS.items():
    _library.impl(_name, _fn, 'CUDA')


def _no_cpu(*args, **kwargs):
    raise _lib.B200PhyError('b200phy ops need CUDA tensors (B200, sm_100a); there is no CPU kernel')


for _name in _IMPLS:
    _library.impl(_name, _no_cpu, 'CPU')

OPS = tuple(_DEFS)
