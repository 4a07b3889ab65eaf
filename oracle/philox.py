"""Shared counter-based random stream (TEST INFRASTRUCTURE — see oracle/__init__.py).

The reference draws from NumPy's global legacy MT19937 (``util/misc.py:354-355``
``randn_c``; ``channels/fading_generators.py:413-425`` Jakes phases;
``apps/awgn_modulators/simulate_psk.py:65`` ``np.random.randint``), which is
sequential and cannot be reproduced in parallel.  The contract of this build
 is instead: host oracle and device kernels consume the SAME
Philox4x32-10 stream, a pure function of ``(seed, stream, unit, word)``:

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (slot & 0xffffffff, (stream << 16) | (slot >> 32) & 0xffff,
               unit & 0xffffffff, unit >> 32)          slot = word // 4
    word w of (stream, unit) = output lane ``w % 4`` of that block

``unit`` is the global realization / frame index, so results do not depend on
batch size or on how units are sharded over GPUs.  Streams: 0 data symbols,
1 channel (Rayleigh H or Jakes phases), 2 noise.

Derived draws (identical formulas in ``pyphysim_b200/csrc/rng.cuh``):
  * data symbol p      = word[p] >> (32 - log2 M)       (uniform on [0, M))
  * uniform  u(x)      = (x + 0.5) * 2**-32  (f64: exact; f32: float32(x)*2**-32
                         + 2**-33 evaluated in float32 — bit-identical on host
                         and device because the multiply is exact)
  * complex normal j   = sqrt(-ln u(word[2j])) * exp(2*pi*i*u(word[2j+1]))
                         (Box-Muller; unit variance, i.e. randn_c's 1/sqrt(2)
                         is folded in: E|c|^2 = 1)
  * Jakes phases       phi_i = 2*pi*u(word[i]), psi_i = 2*pi*u(word[P4 + i]),
                         P4 = P rounded up to a multiple of 4, i in C order of
                         the reference's (L, taps[, Nr, Nt]) phase arrays.
Integers and phases are bit-identical host vs device; the complex normals go
through log/sincos whose last ulp differs, so parity tests either upload the
host draws or read back the device draws (``b200phy_draw_*``).

Known-answer vectors for Philox4x32-10 are the three published with the
Random123 distribution (kat_vectors); ``tests/test_oracle_kat.py`` checks them.
"""
import numpy as np

SEED_DEFAULT = 0x5EEDB200
STREAM_DATA = 0
STREAM_CHANNEL = 1
STREAM_NOISE = 2

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32 with 10 rounds on broadcastable uint32 arrays -> 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1,
                          hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def words(seed, stream, units, n_words, first_word=0):
    """uint32[len(units), n_words]: words first_word.. of (stream, unit)."""
    units = np.atleast_1d(np.asarray(units, dtype=np.uint64))
    assert first_word % 4 == 0
    n_slots = (n_words + 3) // 4
    slots = np.arange(n_slots, dtype=np.uint64) + np.uint64(first_word // 4)
    c0 = (slots & _MASK)[None, :]
    c1 = ((np.uint64(stream) << np.uint64(16)) |
          ((slots >> _S32) & np.uint64(0xFFFF)))[None, :]
    c2 = (units & _MASK)[:, None]
    c3 = (units >> _S32)[:, None]
    o = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.stack(np.broadcast_arrays(*o), axis=-1).reshape(units.size, n_slots * 4)
    return out[:, :n_words]


def uniform(x, dtype=np.float64):
    """(x + 0.5) * 2**-32 in the requested arithmetic (see module docstring)."""
    if np.dtype(dtype) == np.float32:
        return (x.astype(np.float32) * np.float32(2.0 ** -32)) + np.float32(2.0 ** -33)
    return (x.astype(np.float64) + 0.5) * 2.0 ** -32


def data_indices(seed, units, n_symbols, bits):
    """int64[len(units), n_symbols] data symbols, uniform on [0, 2**bits)."""
    w = words(seed, STREAM_DATA, units, n_symbols)
    return (w >> np.uint32(32 - bits)).astype(np.int64)


def cnormal_from_words(w, dtype=np.float64):
    """Complex normals (E|c|^2=1) from word pairs; w[..., 2n] -> [..., n]."""
    u1 = uniform(w[..., 0::2], dtype)
    u2 = uniform(w[..., 1::2], dtype)
    rad = np.sqrt(-np.log(u1))
    ang = (dtype(2.0) * dtype(np.pi)) * u2 if dtype is not np.float64 else 2.0 * np.pi * u2
    cdt = np.complex64 if np.dtype(dtype) == np.float32 else np.complex128
    return (rad * np.cos(ang) + 1j * (rad * np.sin(ang))).astype(cdt)


def cnormal(seed, stream, units, n, dtype=np.float64, first=0):
    """complex[len(units), n]: complex normals first..first+n of (stream, unit)."""
    assert first % 2 == 0
    w = words(seed, stream, units, 2 * n + (2 * n) % 4, first_word=2 * first)
    return cnormal_from_words(w[:, :2 * n], dtype)


def noise_rows(seed, units, n_rows, row_len, dtype=np.float64):
    """Noise draws laid out [unit, row, row_len]; every row starts on a fresh
    Philox slot: complex normal m of row r is normal ``r*2*ceil(row_len/2) + m``."""
    per_row = 2 * ((row_len + 1) // 2)
    c = cnormal(seed, STREAM_NOISE, units, n_rows * per_row, dtype)
    return c.reshape(len(np.atleast_1d(units)), n_rows, per_row)[:, :, :row_len]


def jakes_phases(seed, units, shape, dtype=np.float64):
    """(phi, psi), each [len(units), *shape], phi drawn before psi as the
    reference does (channels/fading_generators.py:413-425)."""
    P = int(np.prod(shape))
    P4 = (P + 3) // 4 * 4
    w = words(seed, STREAM_CHANNEL, units, P4 + P)
    two_pi = np.float32(2.0 * np.pi) if np.dtype(dtype) == np.float32 else 2.0 * np.pi
    phi = two_pi * uniform(w[:, :P], dtype)
    psi = two_pi * uniform(w[:, P4:P4 + P], dtype)
    U = phi.shape[0]
    return phi.reshape((U,) + tuple(shape)), psi.reshape((U,) + tuple(shape))
