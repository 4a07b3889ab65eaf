"""Reference signals (Zadoff-Chu / SRS / DMRS) and pilot-based channel estimators — NumPy restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  SURVEY.md §8f row next-4.  Reference:
``pyphysim/reference_signals/{zadoffchu,root_sequence,srs,dmrs,channel_estimation}.py`` and
``pyphysim/channel_estimation/estimators.py``.

The base sequences of length 12 and 24 are the phase tables of 3GPP TS 36.211 (Tables 5.5.1.2-1 and
5.5.1.2-2, thirty rows each, entries in {-3, -1, 1, 3}); they are stored here two bits per entry
(code c -> 2c - 3) and were checked row by row against the tables the reference carries
(``root_sequence.py:32-63`` and ``:67-218``) by ``tests/golden/make_golden_refsig.py``.
"""
import numpy as np

# one 24-bit (length 12) / 48-bit (length 24) little-endian word per root index: entry n = 2*((w >> 2n) & 3) - 3
PHI12 = (0xcbaf39, 0xc827fa, 0x62040a, 0x7206a9, 0xe6466d, 0x8d6972, 0x8f6c0d, 0xb27254, 0x9e95b2, 0xaa87d2,
         0x70429d, 0xfb8f5b, 0x80a8a2, 0xf1e8cf, 0x9fed18, 0x1ba527, 0x75fe6e, 0x7b0ce8, 0x2508ac, 0x44d6ed,
         0x49baa1, 0x18009d, 0xc8d00a, 0x9e611a, 0x819fba, 0xd4bef2, 0x1d630e, 0x0e6c44, 0x5f3dcd, 0x6cd143)
PHI24 = (0x1339acece72d, 0x2b836ea7080c, 0xf9469dbfcaf7, 0x44adbb94a3a1, 0x520619dfa415, 0xa031d688b9e8,
         0x1d91a658c35a, 0x94e6daeed17c, 0x2020255cc638, 0xa298a7f347ca, 0xd9f981017709, 0x89c7cf01b83e,
         0x44c7ca725abe, 0xc78a7f9bd157, 0xeb485faecbb0, 0x40506ed18e25, 0x4a226fb29571, 0x5747a7f187de,
         0x2b0a48e876aa, 0x0de4663f71be, 0x63674ec45031, 0x24f5992d99a0, 0x1f8b933046c4, 0xf6cf778fbf55,
         0xd16a977531f6, 0x50a38a16b766, 0xa898cb314ae4, 0x444d0b547af1, 0x5dac22865251, 0xe583b6e7745a)


def phi_table(size, root_index):
    """Row ``root_index`` of TS 36.211 Table 5.5.1.2-1 (size 12) / -2 (size 24) as integers in {-3,-1,1,3}
    (root_sequence.py:32-218)."""
    w = (PHI12 if size == 12 else PHI24)[root_index]
    return np.array([2 * ((w >> (2 * n)) & 3) - 3 for n in range(size)], dtype=np.int64)


# primes up to 1009: every Nzc the reference can pick (root_sequence.py:16-28)
def _primes_upto(n):
    s = np.ones(n + 1, dtype=bool)
    s[:2] = False
    for i in range(2, int(n ** 0.5) + 1):
        if s[i]:
            s[i * i::i] = False
    return np.nonzero(s)[0]


PRIMES = _primes_upto(1009)


def largest_prime_leq(n):
    """RootSequence._get_largest_prime_lower_than_number (root_sequence.py:289-305): despite its name it is
    the largest prime <= n from the table of primes <= 1009."""
    return int(PRIMES[PRIMES <= n][-1])


def zc_base(Nzc, u, q=0):
    """calcBaseZC (zadoffchu.py:11-36): a_u[n] = exp(-j pi u n (n + 1 + 2q) / Nzc), u < Nzc."""
    assert u < Nzc
    n = np.arange(Nzc)
    return np.exp((-1j * np.pi * u * n * (n + 1 + 2 * q)) / Nzc)


def shift_seq(root_seq, n_cs, denominator):
    """get_shifted_root_seq (zadoffchu.py:39-72): r[n] * exp(j 2 pi n_cs n / denominator)."""
    assert 0 <= abs(n_cs) < denominator
    n = np.arange(root_seq.size)
    return np.exp(1j * (2 * np.pi * n_cs / denominator) * n) * root_seq


def extend_seq(root_seq, size):
    """get_extended_ZF (zadoffchu.py:75-113): cyclic extension to ``size`` elements."""
    reps = -(-size // root_seq.size)
    return np.tile(root_seq, reps)[:size]


def root_sequence(root_index, size=None, Nzc=None):
    """RootSequence.__init__ / seq_array (root_sequence.py:246-287, 358-370).  Returns (seq, Nzc)."""
    if size is None and Nzc is None:
        raise AttributeError("Either 'size' or 'Nzc' (or both) must be provided.")
    if size is None:
        size = Nzc
    if Nzc is None:
        Nzc = largest_prime_leq(size)
    if size < Nzc:
        raise AttributeError("If 'size' and Nzc are provided, then size must be greater than Nzc")
    if size > 24:
        base = zc_base(Nzc, root_index)
        return (extend_seq(base, size) if size > Nzc else base), Nzc
    if size not in (12, 24):
        raise AttributeError("Invalid root sequence size")
    return np.exp(1j * (np.pi / 4.0) * phi_table(size, root_index)), size


def ue_sequence(root_seq, n_cs, denominator, cover_code=None, normalize=False):
    """SrsUeSequence (srs.py:265-286; denominator 8), DmrsUeSequence (dmrs.py:44-83; denominator 12, optional
    orthogonal cover code as an extra leading axis) with UeSequence's normalisation (srs.py:71-93):
    divide by the norm of the (first cover row of the) sequence."""
    seq = shift_seq(root_seq, n_cs, denominator)
    if cover_code is not None:
        seq = seq * np.asarray(cover_code)[:, None]
    if normalize:
        seq = seq / np.linalg.norm(seq if seq.ndim == 1 else seq[0])
    return seq


def cazac_estimate(ref_seq, received, num_taps_to_keep, size_multiplier=2, normalized=False):
    """CazacBasedChannelEstimator.estimate_channel_freq_domain (reference_signals/channel_estimation.py:69-131):
    y = ifft(conj(r) * Y, Nsc); keep taps 0..num_taps_to_keep; H = fft(taps, size_multiplier * Nsc)
    (times Nsc when the reference sequence was normalised).  ``received`` is [Nsc] or [Nr, Nsc]."""
    r = np.asarray(ref_seq)
    Y = np.asarray(received)
    if Y.ndim not in (1, 2):
        raise ValueError("received_signal must have either one dimension (one receive antenna) or two dimensions")
    y = np.fft.ifft(np.conj(r) * Y, r.size)
    th = y[..., 0:num_taps_to_keep + 1]
    H = np.fft.fft(th, size_multiplier * r.size)
    if normalized:
        H = H * r.size
    return H


def cazac_occ_estimate(ue_seq, cover_code, received, num_taps_to_keep, extra_dimension=True, normalized=False):
    """CazacBasedWithOCCChannelEstimator (channel_estimation.py:134-251): the reference sequence is
    ``ue_seq[0] * cover_code[0]``, the received signal is averaged over the cover-code axis after multiplying
    by the cover code, then the base estimator runs with size_multiplier = 1."""
    cc = np.asarray(cover_code)
    r = np.asarray(received)
    if not extra_dimension:
        if r.ndim == 1:
            r = r.reshape(cc.size, -1)
        elif r.ndim == 2:
            r = r.reshape(r.shape[0], cc.size, -1)
        else:
            raise RuntimeError('Invalid dimension for received_signal: {0}'.format(r.ndim))
    if r.ndim == 2:
        r_mean = np.mean(r * cc[:, None], axis=0)
    elif r.ndim == 3:
        r_mean = np.mean(r * cc[None, :, None], axis=1)
    else:
        raise RuntimeError('Invalid dimension for received_signal: {0}'.format(r.ndim))
    return cazac_estimate(np.asarray(ue_seq)[0] * cc[0], r_mean, num_taps_to_keep, 1, normalized)


def ls_estimate(Y_p, s):
    """compute_ls_estimation (channel_estimation/estimators.py:12-61): Y s^H (s s^H)^-1, batched over a
    leading realization axis of Y_p (and optionally of s)."""
    Y_p, s = np.asarray(Y_p), np.asarray(s)
    if Y_p.ndim == 2:
        return Y_p @ s.T.conj() @ np.linalg.inv(s @ s.conj().T)
    return np.stack([ls_estimate(Y_p[i], s if s.ndim == 2 else s[i]) for i in range(Y_p.shape[0])])


def ls_mse_theory(Nr, noise_power, alpha, pilot_power, num_pilots):
    """compute_theoretical_ls_MSE (estimators.py:64-97)."""
    return Nr * noise_power / ((alpha ** 2) * pilot_power * num_pilots)


def mmse_estimate(Y_p, s, noise_power, C):
    """compute_mmse_estimation (estimators.py:100-174), single tx antenna:
    inv(noise I + P C) C S^H vec(Y) / (s s^H) * P with S = kron(s^T, I_Nr), i.e. S^H vec(Y) = Y s^H."""
    Y_p, s, C = np.asarray(Y_p), np.asarray(s), np.asarray(C)
    if Y_p.ndim == 2:
        assert s.ndim == 2 and s.shape[0] == 1
        Nr, P = Y_p.shape
        W = np.linalg.inv(noise_power * np.eye(Nr) + P * C) @ C
        return W @ (Y_p @ s.T.conj()) / (s @ s.T.conj()) * P
    return np.stack([mmse_estimate(Y_p[i], s if s.ndim == 2 else s[i], noise_power, C) for i in range(Y_p.shape[0])])


def mmse_mse_theory(Nr, noise_power, alpha, pilot_power, num_pilots, C):
    """compute_theoretical_mmse_MSE (estimators.py:177-213)."""
    return np.trace(C @ np.linalg.inv(np.eye(Nr) + alpha ** 2 * pilot_power * num_pilots / noise_power * C))
