"""SURVEY.md §8f row next-4 — reference signals (Zadoff-Chu / SRS / DMRS) and pilot-based channel estimators.

CPU part: the oracle restatement (oracle/refsig.py) against the fixture produced by the unmodified reference
(tests/golden/make_golden_refsig.py -> refsig.npz) and against the known answers the reference's own tests hold.
GPU part: the façade classes (pyphysim_b200.reference_signals / .channel_estimation, through the C ABI) against the
same fixture and the oracle, in the reference's complex128 and in complex64."""
import numpy as np
import pytest

from oracle import philox
from oracle import refsig as O

SEED = 0xC0FFEE


@pytest.fixture(scope='module')
def g(golden):
    return golden('refsig')


def _root_args(row):
    r, s, z = (int(v) for v in row)
    return r, (None if s < 0 else s), (None if z < 0 else z)


# ------------------------------------------------------------------ oracle vs reference fixture (CPU)
def test_oracle_known_answers():
    # tests/reference_signals_package_test.py:99-129 (get_extended_ZF) and the doctest of zadoffchu.py:94-96
    np.testing.assert_array_equal(O.extend_seq(np.array([1, 2, 3, 4, 5]), 8), [1, 2, 3, 4, 5, 1, 2, 3])
    np.testing.assert_array_equal(O.extend_seq(np.arange(1, 4), 10), [1, 2, 3, 1, 2, 3, 1, 2, 3, 1])
    # calcBaseZC definition (reference_signals_package_test.py:68-80): exp(-1j pi u n (n+1) / Nzc)
    n = np.arange(139)
    np.testing.assert_allclose(O.zc_base(139, 25), np.exp(-1j * np.pi * 25 * n * (n + 1) / 139), atol=1e-12)
    # RootSequence sizes (reference_signals_package_test.py:160-181): Nzc = largest prime <= size
    assert [O.largest_prime_leq(v) for v in (64, 150, 300, 1200)] == [61, 149, 293, 1009]
    # the two-PRB tables have unit modulus and phases on the pi/4 grid
    for size in (12, 24):
        for i in range(30):
            t = O.phi_table(size, i)
            assert set(np.unique(t)) <= {-3, -1, 1, 3} and t.size == size
    with pytest.raises(AttributeError):
        O.root_sequence(3)
    with pytest.raises(AttributeError):
        O.root_sequence(3, size=20)


def test_oracle_sequences_vs_reference(g):
    for k, row in enumerate(g['roots']):
        seq, Nzc = O.root_sequence(*_root_args(row))
        np.testing.assert_allclose(seq, g['root%d_seq' % k], atol=2e-10)
        assert seq.size == g['root%d_meta' % k][1] and Nzc == g['root%d_meta' % k][0]
    np.testing.assert_allclose(O.zc_base(31, 5, q=2), g['zc_q'], atol=1e-12)
    np.testing.assert_allclose(O.zc_base(31, 5, q=0.25 + 0.01j), g['zc_qc'], atol=1e-12)
    for k, (rk, ncs, nz) in enumerate(g['srs_users']):
        root, _ = O.root_sequence(*_root_args(g['roots'][rk]))
        np.testing.assert_allclose(O.ue_sequence(root, int(ncs), 8, normalize=bool(nz)), g['srs%d' % k], atol=2e-10)
    for k, (rk, ncs, has_cc, nz) in enumerate(g['dmrs_users']):
        root, _ = O.root_sequence(*_root_args(g['roots'][rk]))
        cc = g['dmrs%d_cc' % k] if has_cc else None
        np.testing.assert_allclose(O.ue_sequence(root, int(ncs), 12, cover_code=cc, normalize=bool(nz)),
                                   g['dmrs%d' % k], atol=2e-10)


def test_oracle_estimators_vs_reference(g):
    for k in range(5):
        mult, keep, nz = (int(v) for v in g['cz%d_par' % k])
        H = O.cazac_estimate(g['cz%d_ref' % k], g['cz%d_Y' % k], keep, mult, bool(nz))
        np.testing.assert_allclose(H, g['cz%d_H' % k], atol=1e-12)
    np.testing.assert_allclose(O.cazac_estimate(g['cz1_ref'], g['cz1_Y'], 8, 3), g['cz5_H'], atol=1e-12)
    for k in range(4):
        keep, extra, nz = (int(v) for v in g['occ%d_par' % k])
        H = O.cazac_occ_estimate(g['occ%d_seq' % k], g['occ%d_cc' % k], g['occ%d_Y' % k], keep, bool(extra), bool(nz))
        np.testing.assert_allclose(H, g['occ%d_H' % k], atol=1e-12)
    for k in range(len(g['ls_cases'])):
        np.testing.assert_allclose(O.ls_estimate(g['ls%d_Y' % k], g['ls%d_s' % k]), g['ls%d_H' % k], atol=1e-12)
    for k in range(len(g['mmse_cases'])):
        H = O.mmse_estimate(g['mmse%d_Y' % k], g['mmse%d_s' % k], 0.5, g['mmse%d_C' % k])
        np.testing.assert_allclose(H, g['mmse%d_H' % k], atol=1e-12)
    np.testing.assert_allclose(O.ls_mse_theory(3, 0.5, 0.7, 1.5, 10), g['ls_mse'])
    np.testing.assert_allclose(O.mmse_mse_theory(3, 0.5, 0.7, 1.5, 10, g['mmse0_C']), g['mmse_mse'])


def test_oracle_ls_single_antenna_is_mean_ratio():
    # tests/channel_estimation_package_test.py:51-76: one tx antenna -> mean(Y / s)
    s = philox.cnormal(SEED, 0, [1], 10).reshape(1, 10)
    Y = philox.cnormal(SEED, 1, [1], 3).reshape(3, 1) @ s + 0.1 * philox.cnormal(SEED, 2, [1], 30).reshape(3, 10)
    ref = (Y @ s.conj().T / (s @ s.conj().T))
    np.testing.assert_allclose(O.ls_estimate(Y, s), ref, atol=1e-13)


# ------------------------------------------------------------------ product (GPU, through the C ABI)
gpu = pytest.mark.gpu


@gpu
def test_gpu_sequences_vs_reference(g):
    from pyphysim_b200.reference_signals.dmrs import DmrsUeSequence, get_dmrs_seq
    from pyphysim_b200.reference_signals.root_sequence import RootSequence
    from pyphysim_b200.reference_signals.srs import SrsUeSequence, get_srs_seq
    from pyphysim_b200.reference_signals.zadoffchu import calcBaseZC, get_extended_ZF, get_shifted_root_seq
    roots = []
    for k, row in enumerate(g['roots']):
        r, s, z = _root_args(row)
        obj = RootSequence(root_index=r, size=s, Nzc=z)
        roots.append(obj)
        assert obj.seq_array().dtype == np.complex128
        np.testing.assert_allclose(obj.seq_array(), g['root%d_seq' % k], atol=2e-10)
        assert [obj.Nzc, obj.size, obj.index] == list(g['root%d_meta' % k])
        np.testing.assert_allclose(obj[3:7], g['root%d_seq' % k][3:7], atol=2e-10)
    np.testing.assert_allclose(calcBaseZC(31, 5, q=2), g['zc_q'], atol=1e-12)
    np.testing.assert_allclose(calcBaseZC(31, 5, q=0.25 + 0.01j), g['zc_qc'], atol=1e-12)
    for k, (rk, ncs, nz) in enumerate(g['srs_users']):
        u = SrsUeSequence(roots[rk], int(ncs), normalize=bool(nz))
        np.testing.assert_allclose(u.seq_array(), g['srs%d' % k], atol=2e-10)
        assert u.normalized == bool(nz) and u.size == g['srs%d' % k].size and u.shape == g['srs%d' % k].shape
    for k, (rk, ncs, has_cc, nz) in enumerate(g['dmrs_users']):
        cc = np.array(g['dmrs%d_cc' % k]) if has_cc else None
        u = DmrsUeSequence(roots[rk], int(ncs), cover_code=cc, normalize=bool(nz))
        np.testing.assert_allclose(u.seq_array(), g['dmrs%d' % k], atol=2e-10)
        assert u.size == int(g['dmrs%d_size' % k])
    # module functions on arbitrary arrays (negative shifts included), against the oracle
    a = philox.cnormal(SEED, 0, [5], 37)
    for ncs, den in ((3, 8), (-5, 12), (0, 8), (11, 12)):
        np.testing.assert_allclose(get_shifted_root_seq(a, ncs, den), O.shift_seq(a, ncs, den), atol=1e-12)
    np.testing.assert_allclose(get_srs_seq(a, 2), O.shift_seq(a, 2, 8), atol=1e-12)
    np.testing.assert_allclose(get_dmrs_seq(a, 7), O.shift_seq(a, 7, 12), atol=1e-12)
    np.testing.assert_array_equal(get_extended_ZF(np.array([1, 2, 3, 4, 5]), 8), [1, 2, 3, 4, 5, 1, 2, 3])
    np.testing.assert_array_equal(get_extended_ZF(np.arange(1, 4), 10), [1, 2, 3, 1, 2, 3, 1, 2, 3, 1])
    # errors the reference raises (root_sequence.py:249-262, 283)
    with pytest.raises(AttributeError):
        RootSequence(3)
    with pytest.raises(AttributeError):
        RootSequence(3, size=20)
    with pytest.raises(AttributeError):
        RootSequence(3, size=20, Nzc=31)
    with pytest.raises(AssertionError):
        get_shifted_root_seq(a, 8, 8)


@gpu
def test_gpu_cazac_estimators_vs_reference(g):
    from pyphysim_b200.reference_signals.channel_estimation import (CazacBasedChannelEstimator,
                                                                    CazacBasedWithOCCChannelEstimator)
    from pyphysim_b200.reference_signals.dmrs import DmrsUeSequence
    from pyphysim_b200.reference_signals.root_sequence import RootSequence
    from pyphysim_b200.reference_signals.srs import SrsUeSequence
    root150 = RootSequence(25, 150, 139)
    users = [SrsUeSequence(root150, 1, normalize=True), SrsUeSequence(root150, 1),
             DmrsUeSequence(root150, 2, normalize=True), SrsUeSequence(RootSequence(7, 24), 0),
             SrsUeSequence(RootSequence(1, 1200), 3, normalize=True)]
    for k, u in enumerate(users):
        mult, keep, nz = (int(v) for v in g['cz%d_par' % k])
        np.testing.assert_allclose(u.seq_array(), g['cz%d_ref' % k], atol=2e-10)
        est = CazacBasedChannelEstimator(u, size_multiplier=mult)
        H = est.estimate_channel_freq_domain(g['cz%d_Y' % k], keep)
        assert H.dtype == np.complex128 and H.shape == g['cz%d_H' % k].shape
        np.testing.assert_allclose(H, g['cz%d_H' % k], atol=1e-10)
    est = CazacBasedChannelEstimator(g['cz1_ref'].copy(), size_multiplier=3)
    np.testing.assert_allclose(est.estimate_channel_freq_domain(g['cz1_Y'], 8), g['cz5_H'], atol=1e-10)
    with pytest.raises(ValueError):
        est.estimate_channel_freq_domain(np.zeros((2, 2, 150), dtype=complex), 8)
    # cover codes
    u2 = DmrsUeSequence(root150, 2, cover_code=np.array([1, -1]), normalize=True)
    occ = CazacBasedWithOCCChannelEstimator(u2)
    np.testing.assert_array_equal(occ.cover_code, [1, -1])
    for k in range(4):
        keep, extra, nz = (int(v) for v in g['occ%d_par' % k])
        np.testing.assert_allclose(u2.seq_array(), g['occ%d_seq' % k], atol=2e-10)
        Y = g['occ%d_Y' % k].copy()
        H = occ.estimate_channel_freq_domain(Y, keep, extra_dimension=bool(extra))
        np.testing.assert_array_equal(Y, g['occ%d_Y' % k])         # the caller's array keeps its shape and values
        np.testing.assert_allclose(H, g['occ%d_H' % k], atol=1e-10)


@gpu
def test_gpu_cazac_batch_f32_and_f64():
    """A batch of realizations x antennas in one launch, CUDA tensors in -> CUDA tensors out, both dtypes."""
    import torch
    from pyphysim_b200.reference_signals.channel_estimation import CazacBasedChannelEstimator
    from pyphysim_b200.reference_signals.root_sequence import RootSequence
    from pyphysim_b200.reference_signals.srs import SrsUeSequence
    u = SrsUeSequence(RootSequence(11, 300), 5, normalize=True)
    n, Nr = 200, 4
    Y = philox.cnormal(SEED, 2, np.arange(n), Nr * 300).reshape(n, Nr, 300)
    ref = O.cazac_estimate(u.seq_array(), Y.reshape(-1, 300), 20, 2, True).reshape(n, Nr, 600)
    est = CazacBasedChannelEstimator(u)
    for cdt, tol in ((torch.complex128, 1e-10), (torch.complex64, 2e-5)):
        H = est.estimate_batch(torch.from_numpy(Y).to(cdt).cuda(), 20)
        assert H.is_cuda and H.dtype == cdt and tuple(H.shape) == (n, Nr, 600)
        err = np.abs(H.cpu().numpy() - ref).max() / np.sqrt(np.mean(np.abs(ref) ** 2))
        assert err < tol, err


@gpu
def test_gpu_pilot_estimators_vs_reference(g):
    from pyphysim_b200.channel_estimation import (compute_ls_estimation, compute_mmse_estimation,
                                                  compute_theoretical_ls_MSE, compute_theoretical_mmse_MSE)
    for k in range(len(g['ls_cases'])):
        H = compute_ls_estimation(g['ls%d_Y' % k], g['ls%d_s' % k])
        assert H.shape == g['ls%d_H' % k].shape and H.dtype == np.complex128
        np.testing.assert_allclose(H, g['ls%d_H' % k], atol=1e-10)
    for k in range(len(g['mmse_cases'])):
        H = compute_mmse_estimation(g['mmse%d_Y' % k], g['mmse%d_s' % k], 0.5, g['mmse%d_C' % k])
        assert H.shape == g['mmse%d_H' % k].shape
        np.testing.assert_allclose(H, g['mmse%d_H' % k], atol=1e-10)
    np.testing.assert_allclose(compute_theoretical_ls_MSE(3, 0.5, 0.7, 1.5, 10), g['ls_mse'])
    np.testing.assert_allclose(compute_theoretical_mmse_MSE(3, 0.5, 0.7, 1.5, 10, g['mmse0_C']), g['mmse_mse'])


@gpu
def test_gpu_pilot_estimators_large_batch_and_mse():
    """1e5 realizations in one launch (the reference loops in Python): LS and MMSE agree with the oracle on a
    sample, and the empirical MSEs match the closed forms (channel_estimation_package_test.py:246-329)."""
    import torch
    from pyphysim_b200.channel_estimation import compute_ls_estimation, compute_mmse_estimation
    n, Nr, P = 100000, 3, 10
    alpha, pp, nv = 0.7, 1.5, 0.5
    units = np.arange(n)
    h = alpha * philox.cnormal(SEED, 1, units, Nr).reshape(n, Nr, 1)
    s = np.sqrt(pp) * np.exp(2j * np.pi * philox.uniform(philox.words(SEED, 0, [7], P))).reshape(1, P)
    N = np.sqrt(nv) * philox.cnormal(SEED, 2, units, Nr * P).reshape(n, Nr, P)
    Y = h @ s + N
    C = alpha ** 2 * np.eye(Nr)
    for cdt, tol in ((torch.complex128, 1e-11), (torch.complex64, 2e-5)):
        Yd, sd = torch.from_numpy(Y).to(cdt).cuda(), torch.from_numpy(s).to(cdt).cuda()
        ls = compute_ls_estimation(Yd, sd).cpu().numpy()
        mm = compute_mmse_estimation(Yd, sd, nv, C).cpu().numpy()
        np.testing.assert_allclose(ls[:500], O.ls_estimate(Y[:500], s), atol=tol, rtol=0)
        np.testing.assert_allclose(mm[:500], O.mmse_estimate(Y[:500], s, nv, C), atol=tol, rtol=0)
    mse_ls = np.mean(np.sum(np.abs(ls - h) ** 2, axis=(1, 2)))
    mse_mm = np.mean(np.sum(np.abs(mm - h) ** 2, axis=(1, 2)))
    np.testing.assert_allclose(mse_ls, O.ls_mse_theory(Nr, nv, 1.0, pp, P), rtol=0.02)
    np.testing.assert_allclose(mse_mm, O.mmse_mse_theory(Nr, nv, 1.0, pp, P, C), rtol=0.02)
    assert mse_mm < mse_ls


@gpu
def test_gpu_estimator_errors():
    from pyphysim_b200.channel_estimation import compute_ls_estimation
    with pytest.raises(ValueError):                                 # fewer pilots than tx antennas: s s^H singular
        compute_ls_estimation(np.zeros((2, 1), dtype=complex), np.ones((2, 1), dtype=complex))
    with pytest.raises(NotImplementedError):
        compute_ls_estimation(np.zeros((2, 8), dtype=complex), np.ones((5, 8), dtype=complex))
