#!/usr/bin/env python
"""Golden fixture for SURVEY.md §8f row next-4 — Zadoff-Chu / SRS / DMRS sequences, the CAZAC-based channel
estimators and the LS / MMSE pilot estimators — produced by the unmodified reference in the build container:
    python tests/golden/make_golden_refsig.py
Channels, pilots and noise come from the oracle's Philox streams, so every case can be regenerated from its seed.
Also asserts that the packed 3GPP phase tables and the prime table of oracle/refsig.py equal the reference's."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')

from pyphysim.channel_estimation import estimators as E  # noqa: E402
from pyphysim.reference_signals import root_sequence as RS  # noqa: E402
from pyphysim.reference_signals.channel_estimation import (CazacBasedChannelEstimator,  # noqa: E402
                                                           CazacBasedWithOCCChannelEstimator)
from pyphysim.reference_signals.dmrs import DmrsUeSequence  # noqa: E402
from pyphysim.reference_signals.srs import SrsUeSequence  # noqa: E402
from pyphysim.reference_signals.zadoffchu import calcBaseZC  # noqa: E402

from make_golden import SEED  # noqa: E402
from oracle import philox, refsig  # noqa: E402

for i in range(30):
    assert np.array_equal(refsig.phi_table(12, i), RS.ROOT_TABLE1[str(i)])
    assert np.array_equal(refsig.phi_table(24, i), RS.ROOT_TABLE2[str(i)])
assert np.array_equal(refsig.PRIMES, RS._SMALL_PRIME_LIST)

out = {}


def cn(stream, unit, n):
    return philox.cnormal(SEED, stream, [unit], n)[0]


# ---- root sequences: (root_index, size, Nzc)
ROOTS = [(25, None, 139), (25, 150, 139), (6, 64, None), (17, 300, None), (12, 12, None), (7, 24, None),
         (29, 24, None), (3, 48, 31), (1, 1200, None), (22, 63, 31)]
out['roots'] = np.array([[r, -1 if s is None else s, -1 if z is None else z] for r, s, z in ROOTS])
for k, (r, s, z) in enumerate(ROOTS):
    obj = RS.RootSequence(root_index=r, size=s, Nzc=z)
    out['root%d_seq' % k] = obj.seq_array()
    out['root%d_meta' % k] = np.array([obj.Nzc, obj.size, obj.index])
out['zc_q'] = calcBaseZC(31, 5, q=2)
out['zc_qc'] = calcBaseZC(31, 5, q=0.25 + 0.01j)

# ---- user sequences
USERS = [(1, 3, False), (1, 7, True), (3, 0, True), (4, 5, False)]          # (root case, n_cs, normalize)
out['srs_users'] = np.array([[a, b, int(c)] for a, b, c in USERS])
for k, (rk, ncs, nz) in enumerate(USERS):
    r, s, z = ROOTS[rk]
    out['srs%d' % k] = SrsUeSequence(RS.RootSequence(r, s, z), ncs, normalize=nz).seq_array()
DMRS = [(1, 11, None, False), (3, 4, (1, -1), True), (4, 2, (1, 1), False), (5, 9, (1, -1), True)]
out['dmrs_users'] = np.array([[a, b, 0 if c is None else 1, int(d)] for a, b, c, d in DMRS])
for k, (rk, ncs, cc, nz) in enumerate(DMRS):
    r, s, z = ROOTS[rk]
    obj = DmrsUeSequence(RS.RootSequence(r, s, z), ncs, cover_code=None if cc is None else np.array(cc), normalize=nz)
    out['dmrs%d' % k] = obj.seq_array()
    out['dmrs%d_size' % k] = np.array(obj.size)
    if cc is not None:
        out['dmrs%d_cc' % k] = np.array(cc)


# ---- CAZAC estimators: two users on the same resource, sparse random channels
def chan_freq(unit, nsc, n_taps=12, nr=1):
    taps = cn(1, unit, nr * n_taps).reshape(nr, n_taps) * np.exp(-0.2 * np.arange(n_taps))
    return np.fft.fft(taps, nsc, axis=1)


def cazac_case(tag, seq1, seq2, mult, nr, unit, keep):
    n = seq1.size
    nsc = mult * n
    H1, H2 = chan_freq(unit, nsc, nr=nr), chan_freq(unit + 1, nsc, nr=nr)
    comb = np.arange(0, nsc, mult)
    Y = H1[:, comb] * seq1.seq_array() + H2[:, comb] * seq2.seq_array() + 0.05 * cn(2, unit, nr * n).reshape(nr, n)
    if nr == 1:
        Y = Y[0]
    est = CazacBasedChannelEstimator(seq1, size_multiplier=mult)
    out[tag + '_Y'] = Y
    out[tag + '_ref'] = seq1.seq_array()
    out[tag + '_par'] = np.array([mult, keep, int(seq1.normalized)])
    out[tag + '_H'] = est.estimate_channel_freq_domain(Y, keep)


root150 = RS.RootSequence(25, 150, 139)
cazac_case('cz0', SrsUeSequence(root150, 1, normalize=True), SrsUeSequence(root150, 4, normalize=True), 2, 1, 900, 15)
cazac_case('cz1', SrsUeSequence(root150, 1), SrsUeSequence(root150, 4), 1, 1, 910, 15)
cazac_case('cz2', DmrsUeSequence(root150, 2, normalize=True), DmrsUeSequence(root150, 7, normalize=True), 1, 3, 920, 11)
root24 = RS.RootSequence(7, 24)
cazac_case('cz3', SrsUeSequence(root24, 0), SrsUeSequence(root24, 5), 2, 2, 930, 3)
root1200 = RS.RootSequence(1, 1200)
cazac_case('cz4', SrsUeSequence(root1200, 3, normalize=True), SrsUeSequence(root1200, 6, normalize=True), 1, 4, 940, 30)
# a plain ndarray as the reference sequence (channel_estimation.py:56-60)
est = CazacBasedChannelEstimator(out['cz1_ref'].copy(), size_multiplier=3)
out['cz5_H'] = est.estimate_channel_freq_domain(out['cz1_Y'], 8)

# with orthogonal cover codes: two users, same shift family, cover codes (1, 1) and (1, -1)
for k, (nr, extra) in enumerate([(1, True), (3, True), (1, False), (2, False)]):
    cc1, cc2 = np.array([1, 1]), np.array([1, -1])
    u1 = DmrsUeSequence(root150, 2, cover_code=cc1, normalize=True)
    u2 = DmrsUeSequence(root150, 2, cover_code=cc2, normalize=True)
    n = u1.size
    H1, H2 = chan_freq(950 + 2 * k, n, nr=nr), chan_freq(951 + 2 * k, n, nr=nr)
    Y = (H1[:, None, :] * u1.seq_array()[None] + H2[:, None, :] * u2.seq_array()[None]
         + 0.05 * cn(2, 950 + k, nr * 2 * n).reshape(nr, 2, n))
    if nr == 1:
        Y = Y[0]
    if not extra:
        Y = Y.reshape(n * 2) if nr == 1 else Y.reshape(nr, 2 * n)
    tag = 'occ%d' % k
    out[tag + '_Y'] = Y
    out[tag + '_seq'] = u2.seq_array()
    out[tag + '_cc'] = cc2
    out[tag + '_par'] = np.array([9, int(extra), 1])
    out[tag + '_H'] = CazacBasedWithOCCChannelEstimator(u2).estimate_channel_freq_domain(Y.copy(), 9, extra_dimension=extra)

# ---- LS / MMSE pilot estimators
LS = [(1, 3, 1, 10, False), (1, 3, 2, 10, False), (5, 3, 2, 10, False), (5, 4, 2, 16, True), (7, 2, 4, 12, True),
      (4, 8, 3, 9, False)]                               # (realizations, Nr, Nt, pilots, per-realization pilots)
out['ls_cases'] = np.array([[a, b, c, d, int(e)] for a, b, c, d, e in LS])
for k, (nre, nr, nt, P, per) in enumerate(LS):
    H = 0.7 * cn(1, 1000 + k, nre * nr * nt).reshape(nre, nr, nt)
    s = np.sqrt(1.5) * cn(0, 1000 + k, (nre if per else 1) * nt * P).reshape(-1, nt, P)
    N = np.sqrt(0.5) * cn(2, 1000 + k, nre * nr * P).reshape(nre, nr, P)
    Y = H @ s + N
    if nre == 1:
        Y, s_in = Y[0], s[0]
    else:
        s_in = s if per else s[0]
    out['ls%d_Y' % k], out['ls%d_s' % k] = Y, s_in
    out['ls%d_H' % k] = E.compute_ls_estimation(Y, s_in)
MM = [(1, 3, 10, False), (6, 3, 10, False), (6, 4, 20, True), (3, 8, 7, True)]
out['mmse_cases'] = np.array([[a, b, c, int(d)] for a, b, c, d in MM])
for k, (nre, nr, P, per) in enumerate(MM):
    A = cn(1, 1100 + k, nr * nr).reshape(nr, nr)
    C = A @ A.conj().T / nr + 0.1 * np.eye(nr)             # a Hermitian positive definite covariance
    h = cn(1, 1110 + k, nre * nr).reshape(nre, nr, 1)
    s = np.sqrt(1.5) * cn(0, 1100 + k, (nre if per else 1) * P).reshape(-1, 1, P)
    N = np.sqrt(0.5) * cn(2, 1100 + k, nre * nr * P).reshape(nre, nr, P)
    Y = h @ s + N
    if nre == 1:
        Y, s_in = Y[0], s[0]
    else:
        s_in = s if per else s[0]
    out['mmse%d_Y' % k], out['mmse%d_s' % k], out['mmse%d_C' % k] = Y, s_in, C
    out['mmse%d_H' % k] = E.compute_mmse_estimation(Y, s_in, 0.5, C)
out['ls_mse'] = np.array(E.compute_theoretical_ls_MSE(3, 0.5, 0.7, 1.5, 10))
out['mmse_mse'] = np.array(E.compute_theoretical_mmse_MSE(3, 0.5, 0.7, 1.5, 10, out['mmse0_C']))

np.savez_compressed(os.path.join(HERE, 'refsig.npz'), **out)
print('wrote refsig.npz: %d arrays, %d bytes' % (len(out), os.path.getsize(os.path.join(HERE, 'refsig.npz'))))
