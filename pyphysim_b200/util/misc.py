"""Mirror of the hot-path helpers of pyphysim/util/misc.py; array arguments run on the GPU."""
import itertools
import math

import numpy as np

from .. import _device as D
from .. import _lib

_randn_counter = itertools.count()
_STREAM_MISC = 3          # Philox stream id used by the stand-alone randn_c (0..2 are the link streams)


def level2bits(n):
    """util/misc.py:392-414."""
    if n < 1:
        raise ValueError("level2bits: n must be greater then one")
    return int2bits(n - 1)


def int2bits(n):
    """util/misc.py:417-446."""
    if n < 0:
        raise ValueError("int2bits: n must be greater then zero")
    if n == 0:
        return 1
    return int(n).bit_length()


def xor(a, b):
    """util/misc.py:297-324."""
    return a ^ b


def count_bits(n):
    """util/misc.py:449-476: elementwise popcount (b200phy_count_bits)."""
    if np.isscalar(n):
        return bin(int(n)).count('1') if n > 0 else 0
    lib = _lib.load()
    a, was_np = D.to_device(n, np.int64)
    torch = _lib.torch_cuda()
    out = torch.empty_like(a)
    _lib.check(lib.b200phy_count_bits(_lib.ptr(a), a.numel(), _lib.ptr(out), _lib.cur_stream()))
    return D.from_device(out, was_np)


def count_bit_errors(first, second, axis=None):
    """util/misc.py:519-566: sum(count_bits(first ^ second), axis)."""
    if np.isscalar(first) and np.isscalar(second):
        return bin(int(first) ^ int(second)).count('1')
    lib = _lib.load()
    torch = _lib.torch_cuda()
    a, was_np = D.to_device(first, np.int64)
    b, _ = D.to_device(second, np.int64)
    if a.shape != b.shape:
        a, b = torch.broadcast_tensors(a, b)
        a, b = a.contiguous(), b.contiguous()
    if axis is None:
        out = torch.zeros(2, dtype=torch.int64, device=a.device)
        _lib.check(lib.b200phy_count_errors(_lib.ptr(a), _lib.ptr(b), a.numel(), _lib.ptr(out),
                                            _lib.cur_stream()))
        return int(out[1].item())
    bits = count_bits(torch.bitwise_xor(a, b))
    res = bits.sum(dim=axis)
    return D.from_device(res, was_np)


def count_symbol_and_bit_errors(first, second):
    """(sum(first != second), count_bit_errors(first, second)) in one device pass."""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    a, _ = D.to_device(first, np.int64)
    b, _ = D.to_device(second, np.int64)
    out = torch.zeros(2, dtype=torch.int64, device=a.device)
    _lib.check(lib.b200phy_count_errors(_lib.ptr(a), _lib.ptr(b), a.numel(), _lib.ptr(out),
                                        _lib.cur_stream()))
    o = out.cpu()
    return int(o[0]), int(o[1])


def randn_c(*args, seed=None, dtype=np.complex128, device_out=False):
    """util/misc.py:327-355: circularly-symmetric complex normals with unit variance.

    The reference draws from NumPy's global MT19937; here the samples come from the shared Philox
    stream (stream 3, one fresh `unit` per call unless `seed` pins it), generated on the GPU."""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    dt = _lib.parse_dtype(dtype)
    shape = tuple(int(a) for a in args)
    n = int(np.prod(shape)) if shape else 1
    x = torch.zeros(n, dtype=_lib.cplx_dtype(dt), device='cuda')
    unit = next(_randn_counter)
    from .. import SEED_DEFAULT
    _lib.check(lib.b200phy_awgn(dt, _lib.ptr(x), n, 1.0, SEED_DEFAULT if seed is None else seed,
                                _STREAM_MISC, unit, 0, _lib.cur_stream()))
    x = x.reshape(shape) if shape else x.reshape(())
    if device_out:
        return x
    out = x.cpu().numpy()
    return out if shape else complex(out)


def gmd(U, S, V_H, tol=0.0):
    """util/misc.py:18-159: geometric mean decomposition A = Q R P^H from an SVD A = U diag(S) V_H, on the
    GPU (``b200phy_gmd``).  Shapes follow the reference: U [m, m] (or thin [m, n]), S [n], V_H [n, n] ->
    Q like U, R [m, n] real upper triangular with constant diagonal, P [n, n].  All singular values must be
    kept (every S >= tol) and n <= m <= 4."""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    U = np.asarray(U, dtype=np.complex128)
    S = np.asarray(S, dtype=np.float64)
    V = np.ascontiguousarray(np.asarray(V_H, dtype=np.complex128).conj().T)
    m, n = U.shape[0], V.shape[0]
    if int(np.sum(S >= tol)) < 1:
        raise RuntimeError("This is no singular value greater than the tolerance")
    if int(np.sum(S >= tol)) != n or S.size != n:
        raise NotImplementedError("gmd on the GPU keeps every singular value (tol drops %d of %d)"
                                  % (n - int(np.sum(S >= tol)), n))
    u, _ = D.to_device(np.ascontiguousarray(U[:, :n]), np.complex128)
    s, _ = D.to_device(S, np.float64)
    v, _ = D.to_device(V, np.complex128)
    q = torch.empty_like(u)
    r = torch.empty((n, n), dtype=torch.float64, device='cuda')
    p = torch.empty_like(v)
    _lib.check(lib.b200phy_gmd(_lib.ptr(u), _lib.ptr(s), _lib.ptr(v), 1, m, n, _lib.ptr(q), _lib.ptr(r),
                               _lib.ptr(p), _lib.cur_stream()))
    Q = U.copy()
    Q[:, :n] = q.cpu().numpy()
    R = np.zeros((m, n))
    R[:n, :] = r.cpu().numpy()
    return Q, R, p.cpu().numpy()


def qfunc(x):
    """util/misc.py:569-592 (host scalar; theory curves only)."""
    return 0.5 * math.erfc(x / math.sqrt(2))


def pretty_time(time_in_seconds):
    """util/misc.py:595-640 (formatting helper used by SimulationRunner)."""
    seconds = time_in_seconds
    minutes = int(seconds) // 60
    seconds = int(round(seconds % 60))
    hours = minutes // 60
    minutes %= 60
    days = hours // 24
    hours %= 24
    out = []
    if days > 0:
        out.append("%sd" % days)
    if hours > 0:
        out.append("%sh" % hours)
    if minutes > 0:
        out.append("%sm" % minutes)
    if seconds > 0 or not out:
        out.append("%ss" % ("%.2f" % time_in_seconds if not out and time_in_seconds < 1 else seconds))
    return ":".join(out)
