"""ctypes binding of libb200phy.so (the C ABI declared in include/b200phy.h).

The product path has no CPU fallback: if the library cannot be loaded, or a kernel is asked to run
without a CUDA device, this module raises.  PyTorch is used only for device memory and streams.
"""
import ctypes as C
import os
import threading

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('B200PHY_LIB') or os.path.join(PKG, 'libb200phy.so')   # override: A/B builds

F32, F64 = 0, 1
MODEM_TABLE, MODEM_QAM, MODEM_BPSK, MODEM_QPSK = 0, 1, 2, 3
JAKES_AUTO, JAKES_RECURRENCE, JAKES_POLY = 0, 1, 2
ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_RANGE = 1, 2, 3, 4
MAX_TAPS, MAX_RAYS, MAX_ANT = 32, 64, 4

_u8p = C.POINTER(C.c_uint8)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p


class Modem(C.Structure):
    _fields_ = [('kind', C.c_int32), ('M', C.c_int32), ('table', C.c_void_p)]


class OfdmTdlParams(C.Structure):
    _fields_ = [('struct_size', C.c_int32), ('dtype', C.c_int32),
                ('fft', C.c_int32), ('cp', C.c_int32), ('used', C.c_int32), ('n_sym', C.c_int32),
                ('Nr', C.c_int32), ('Nt', C.c_int32), ('n_taps', C.c_int32), ('L', C.c_int32),
                ('jakes_mode', C.c_int32), ('reserved', C.c_int32),
                ('delays', C.c_int32 * MAX_TAPS), ('tap_powers', C.c_double * MAX_TAPS),
                ('Fd', C.c_double), ('Ts', C.c_double), ('t0', C.c_double),
                ('noise_var', C.c_double), ('filter_noise_var', C.c_double),
                ('seed', C.c_uint64)]


_MP = C.POINTER(Modem)
_PP = C.POINTER(OfdmTdlParams)

# name -> (restype, argtypes); must list every symbol of include/b200phy.h
SIGNATURES = {
    'b200phy_version': (C.c_int, []),
    'b200phy_last_error': (C.c_char_p, []),
    'b200phy_launch_count': (C.c_uint64, []),
    'b200phy_last_kernel': (C.c_char_p, []),
    'b200phy_map': (C.c_int, [C.c_int, _MP, _vp, C.c_int64, _vp, _vp, _vp]),
    'b200phy_demap': (C.c_int, [C.c_int, _MP, _vp, C.c_int64, _vp, _vp]),
    'b200phy_count_errors': (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp]),
    'b200phy_count_bits': (C.c_int, [_vp, C.c_int64, _vp, _vp]),
    'b200phy_awgn': (C.c_int, [C.c_int, _vp, C.c_int64, C.c_double, C.c_uint64, C.c_uint32,
                               C.c_uint64, C.c_uint64, _vp]),
    'b200phy_ofdm_mod': (C.c_int, [C.c_int, _vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    'b200phy_ofdm_demod': (C.c_int, [C.c_int, _vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    'b200phy_jakes': (C.c_int, [C.c_int, _vp, _vp, C.c_int, C.c_int64, C.c_int64, C.c_double,
                                C.c_double, C.c_double, _vp, _vp]),
    'b200phy_tdl_apply': (C.c_int, [C.c_int, _vp, _vp, _f64p, _i32p, C.c_int, C.c_int, C.c_int,
                                    C.c_int64, _vp, _vp]),
    'b200phy_tdl_freq_response': (C.c_int, [C.c_int, _vp, _i32p, C.c_int, C.c_int64, C.c_int64, C.c_int, _vp, _vp]),
    'b200phy_freq_apply': (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, _vp, _vp]),
    'b200phy_scale_rows': (C.c_int, [C.c_int, _vp, C.c_int, C.c_int64, _f64p, _vp]),
    'b200phy_ofdm_equalize': (C.c_int, [C.c_int, _vp, _vp, _f64p, _i32p, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _vp, _vp]),
    'b200phy_blast_decode': (C.c_int, [C.c_int, _vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int,
                                       C.c_double, _vp, _vp]),
    'b200phy_alamouti_encode': (C.c_int, [C.c_int, _vp, C.c_int64, C.c_int, _vp, _vp]),
    'b200phy_alamouti_decode': (C.c_int, [C.c_int, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, _vp]),
    'b200phy_link_siso_flat': (C.c_int, [C.c_int, _MP, C.c_int, C.c_double, C.c_uint64, C.c_uint64,
                                         C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'b200phy_link_alamouti': (C.c_int, [C.c_int, _MP, C.c_int, C.c_int, C.c_double, C.c_uint64,
                                        C.c_uint64, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'b200phy_link_blast': (C.c_int, [C.c_int, _MP, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                     C.c_uint64, C.c_uint64, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp,
                                     _vp]),
    'b200phy_link_precoded': (C.c_int, [C.c_int, _MP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                        C.c_double, C.c_uint64, C.c_uint64, C.c_int64, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp]),
    'b200phy_svd': (C.c_int, [_vp, C.c_int64, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    'b200phy_gmd': (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    'b200phy_mat_apply': (C.c_int, [C.c_int, _vp, C.c_int, C.c_int, _vp, C.c_int64, _vp, _vp]),
    'b200phy_refsig_sequence': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _vp, C.c_int, C.c_int,
                                          C.c_int, C.c_double, C.c_double, _vp, _vp]),
    'b200phy_cazac_estimate': (C.c_int, [C.c_int, _vp, _vp, _f64p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int,
                                         C.c_double, _vp, _vp]),
    'b200phy_ls_estimate': (C.c_int, [C.c_int, _vp, _vp, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    'b200phy_mmse_estimate': (C.c_int, [C.c_int, _vp, _vp, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_double, _f64p,
                                        _vp, _vp]),
    'b200phy_link_ofdm_tdl': (C.c_int, [_PP, _MP, C.c_uint64, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp]),
    'b200phy_draw_siso_flat': (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int64, _vp, _vp,
                                         _vp, _vp]),
    'b200phy_draw_flat_mimo': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_uint64, C.c_uint64, C.c_int64, _vp, _vp, _vp, _vp]),
    'b200phy_draw_ofdm_tdl': (C.c_int, [_PP, C.c_int, C.c_uint64, C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    'b200phy_link_siso_flat_host': (C.c_int, [C.c_int, C.c_int, C.c_int, _f64p, C.c_int, C.c_double,
                                              C.c_uint64, C.c_uint64, C.c_int64, _vp, _vp, _vp, _vp,
                                              _i64p]),
    'b200phy_link_ofdm_tdl_host': (C.c_int, [_PP, C.c_int, C.c_int, _f64p, C.c_uint64, C.c_int64, _vp,
                                             _vp, _vp, _vp, _vp, _i64p]),
    'b200phy_ofdm_tdl_check_params': (C.c_int, [_PP]),
    'b200phy_link_alamouti_host': (C.c_int, [C.c_int, C.c_int, C.c_int, _f64p, C.c_int, C.c_int, C.c_double,
                                             C.c_uint64, C.c_uint64, C.c_int64, _vp, _vp, _vp, _vp, _i64p]),
    'b200phy_link_blast_host': (C.c_int, [C.c_int, C.c_int, C.c_int, _f64p, C.c_int, C.c_int, C.c_int,
                                          C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_int64, _vp, _vp,
                                          _vp, _vp, _i64p]),
    'b200phy_link_precoded_host': (C.c_int, [C.c_int, C.c_int, C.c_int, _f64p, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_int64, _vp, _vp,
                                             _vp, _vp, _i64p]),
}

_lib = None
_lock = threading.Lock()


class B200PhyError(RuntimeError):
    pass


def load():
    """Load libb200phy.so (building it with nvcc if it is missing).  Raises if neither works."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if 'B200PHY_LIB' not in os.environ:
            # the .so is git-ignored but travels with snapshots: rebuild when csrc/ or include/ changed since it
            # was built (source-digest check, a no-op when current).  Without nvcc (a pure runtime box) the
            # shipped binary is used as is; b200phy_version / struct_size still guard the ABI.
            from . import _build
            if not os.path.exists(LIB_PATH) or _build.have_nvcc():
                _build.build()
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:                       # no silent fallback: this IS the product
            raise B200PhyError('cannot load %s: %s' % (LIB_PATH, e))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)                # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    """Map a B200PHY_ERR_* return code to the exception type the reference raises."""
    if rc == 0:
        return
    msg = load().b200phy_last_error().decode()
    if rc in (ERR_INVALID, ERR_RANGE):
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise B200PhyError(msg)


# ---- torch plumbing (device memory + streams only) ---------------------------------------------
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        raise B200PhyError('pyphysim_b200 needs a CUDA device (B200, sm_100a); none is visible and '
                           'there is no CPU fallback')
    return torch


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def real_dtype(dtype):
    import torch
    return torch.float32 if dtype == F32 else torch.float64


def cplx_dtype(dtype):
    import torch
    return torch.complex64 if dtype == F32 else torch.complex128


def parse_dtype(d):
    if d in (F32, 'f32', 'float32', 'complex64', np.float32, np.complex64):
        return F32
    if d in (F64, 'f64', 'float64', 'complex128', np.float64, np.complex128, complex, float):
        return F64
    import torch
    if d in (torch.float32, torch.complex64):
        return F32
    if d in (torch.float64, torch.complex128):
        return F64
    raise ValueError('unknown dtype %r' % (d,))
