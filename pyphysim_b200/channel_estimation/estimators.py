"""Pilot-based LS / MMSE channel estimators with the API of pyphysim/channel_estimation/estimators.py.  The
reference loops over realizations in Python; here every realization is one warp of one kernel launch
(``b200phy_ls_estimate`` / ``b200phy_mmse_estimate``)."""
import ctypes as C

import numpy as np

from .. import _device as D
from .. import _lib

C_F64P = C.POINTER(C.c_double)

__all__ = ['compute_ls_estimation', 'compute_theoretical_ls_MSE', 'compute_mmse_estimation',
           'compute_theoretical_mmse_MSE']


def _prep(Y_p, s):
    if Y_p.ndim == 2:
        assert s.ndim == 2
    dtype = D.dtype_of_samples(Y_p) if D.is_torch(Y_p) else _lib.F64
    y, was_np = D.to_device(Y_p, D.complex_np(dtype))
    sd, _ = D.to_device(s, D.complex_np(dtype))
    single = y.dim() == 2
    if single:
        y = y.reshape((1,) + tuple(y.shape))
    batch, Nr, P = (int(v) for v in y.shape)
    per_unit = sd.dim() == 3
    if per_unit:
        assert sd.shape[0] == batch
    Nt = int(sd.shape[-2])
    if int(sd.shape[-1]) != P:
        raise ValueError("pilots and received signal disagree on the number of pilot symbols")
    return y, sd, dtype, was_np, single, batch, Nr, Nt, P, per_unit


def compute_ls_estimation(Y_p, s):
    """estimators.py:12-61.  Y_p: [Nr, P] or [realizations, Nr, P]; s: [Nt, P] or [realizations, Nt, P]
    -> [Nr, Nt] or [realizations, Nr, Nt]."""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    y, sd, dtype, was_np, single, batch, Nr, Nt, P, per_unit = _prep(Y_p, s)
    out = torch.empty((batch, Nr, Nt), dtype=_lib.cplx_dtype(dtype), device='cuda')
    _lib.check(lib.b200phy_ls_estimate(dtype, _lib.ptr(y), _lib.ptr(sd), int(per_unit), batch, Nr, Nt, P,
                                       _lib.ptr(out), _lib.cur_stream()))
    return D.from_device(out[0] if single else out, was_np)


def compute_theoretical_ls_MSE(Nr, noise_power, alpha, pilot_power, num_pilots):
    """estimators.py:64-97 (a scalar formula)."""
    return Nr * noise_power / ((alpha ** 2) * pilot_power * num_pilots)


def compute_mmse_estimation(Y_p, s, noise_power, C):
    """estimators.py:100-174 (one transmit antenna).  C: [Nr, Nr] channel covariance -> [Nr, 1] or
    [realizations, Nr, 1]."""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    y, sd, dtype, was_np, single, batch, Nr, Nt, P, per_unit = _prep(Y_p, s)
    assert Nt == 1
    cov = np.ascontiguousarray(np.asarray(C.cpu() if D.is_torch(C) else C, dtype=np.complex128))
    if cov.shape != (Nr, Nr):
        raise ValueError("C must be a %d x %d covariance matrix" % (Nr, Nr))
    flat = cov.view(np.float64).reshape(-1)
    out = torch.empty((batch, Nr, 1), dtype=_lib.cplx_dtype(dtype), device='cuda')
    _lib.check(lib.b200phy_mmse_estimate(dtype, _lib.ptr(y), _lib.ptr(sd), int(per_unit), batch, Nr, P,
                                         float(noise_power), flat.ctypes.data_as(C_F64P), _lib.ptr(out),
                                         _lib.cur_stream()))
    return D.from_device(out[0] if single else out, was_np)


def compute_theoretical_mmse_MSE(Nr, noise_power, alpha, pilot_power, num_pilots, C):
    """estimators.py:177-213: trace(C (I + alpha^2 pilot_power num_pilots / noise_power C)^-1); scalar bookkeeping
    on one Nr x Nr matrix (host)."""
    C = np.asarray(C)
    return np.trace(C @ np.linalg.inv(np.eye(Nr) + alpha ** 2 * pilot_power * num_pilots / noise_power * C))
