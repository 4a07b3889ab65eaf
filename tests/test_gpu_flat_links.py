"""GPU parity of the flat-fading links (configs C2, C4 + Blast) against the oracle and the golden
fixtures.  Everything goes through the C ABI (pyphysim_b200.links -> libb200phy.so)."""
import numpy as np
import pytest

from oracle import links as OL
from oracle import modulators as md
from oracle import philox

from gpu_util import (assert_decisions, assert_samples_close, cuda, oracle_modem, product_modem)

pytestmark = pytest.mark.gpu
SEED = 0xC0FFEE


def _t(x):
    return x.cpu().numpy()


# ------------------------------------------------------------------ RNG contract
def test_device_draws_match_host_philox():
    from pyphysim_b200 import links
    pm = product_modem('qam', 64)
    n = 5000
    for dtype, npdt, tol in (('f64', np.float64, 1e-13), ('f32', np.float32, 2e-6)):
        idx, h, noise = links.draw_siso_flat(pm, n, seed=SEED, first_unit=7, dtype=dtype)
        hi, hh, hn = OL.draws_siso_flat(SEED, np.arange(7, 7 + n), 6, dtype=npdt)
        assert np.array_equal(_t(idx), hi)                       # integers: bit exact
        np.testing.assert_allclose(_t(h), hh, atol=tol, rtol=tol)
        np.testing.assert_allclose(_t(noise), hn, atol=tol, rtol=tol)
    idx, H, noise = links.draw_flat_mimo(pm, 300, Nr=3, Nt=2, num_symbols=5, n_data=10, seed=SEED,
                                         first_unit=11, dtype='f64')
    hi, hH, hn = OL.draws_flat_mimo(SEED, np.arange(11, 311), 6, 3, 2, 5, 10)
    assert np.array_equal(_t(idx), hi)
    np.testing.assert_allclose(_t(H), hH, atol=1e-13)
    np.testing.assert_allclose(_t(noise), hn, atol=1e-13)


# ------------------------------------------------------------------ C2: SISO flat
@pytest.mark.parametrize('kind,M,rayleigh', [('qam', 64, True), ('qam', 16, False), ('qam', 256, True),
                                             ('psk', 8, True), ('qpsk', 4, False), ('bpsk', 2, True)])
def test_siso_flat_f64_bit_exact_vs_oracle(kind, M, rayleigh):
    from pyphysim_b200 import links
    import torch
    pm, om = product_modem(kind, M), oracle_modem(kind, M)
    n, nv = 20003, 1 / md.dB2Linear(12.0)
    units = np.arange(3, 3 + n)
    idx, h, nz = OL.draws_siso_flat(SEED, units, om.bits, rayleigh)
    ref_hat, ref_r = OL.siso_flat(om, idx, h, nz, nv)
    draws = (cuda(idx.astype(np.uint8)), cuda(h) if rayleigh else None, cuda(nz))
    cnt, hat, dec = links.link_siso_flat(pm, nv, n, rayleigh=rayleigh, dtype='f64', draws=draws,
                                         want_idx=True, want_samples=True)
    assert_samples_close(_t(dec), ref_r, 1e-12, 'siso f64 samples')
    assert_decisions(_t(hat), ref_hat, om, ref_r, exact=True, what='siso f64')
    assert np.array_equal(cnt, OL.counters(idx, ref_hat, om.bits))
    assert cnt[0] > 0
    # fused mode = stream mode on the device's own draws, and invariant to how the batch is split
    d_idx, d_h, d_n = links.draw_siso_flat(pm, n, rayleigh=rayleigh, seed=SEED, first_unit=3, dtype='f64')
    c_stream, hat_s = links.link_siso_flat(pm, nv, n, rayleigh=rayleigh, dtype='f64',
                                           draws=(d_idx, d_h, d_n), want_idx=True)
    c_fused, hat_f = links.link_siso_flat(pm, nv, n, rayleigh=rayleigh, dtype='f64', seed=SEED,
                                          first_unit=3, want_idx=True)
    assert np.array_equal(c_stream, c_fused) and torch.equal(hat_s, hat_f)
    acc = torch.zeros(4, dtype=torch.int64, device='cuda')
    for a, b in ((0, 1000), (1000, 7777), (7777, n)):
        links.link_siso_flat(pm, nv, b - a, rayleigh=rayleigh, dtype='f64', seed=SEED,
                             first_unit=3 + a, counters=acc)
    assert np.array_equal(_t(acc), c_fused)


def test_siso_flat_f32_vs_oracle_on_device_draws():
    from pyphysim_b200 import links
    pm, om = product_modem('qam', 64), oracle_modem('qam', 64)
    n, nv = 200000, 1 / md.dB2Linear(15.0)
    d_idx, d_h, d_n = links.draw_siso_flat(pm, n, seed=SEED, first_unit=0, dtype='f32')
    cnt, hat, dec = links.link_siso_flat(pm, nv, n, dtype='f32', seed=SEED, want_idx=True,
                                         want_samples=True)
    idx = _t(d_idx).astype(np.int64)
    ref_hat, ref_r = OL.siso_flat(om, idx, _t(d_h).astype(complex), _t(d_n).astype(complex), nv)
    assert_samples_close(_t(dec), ref_r, 1e-5, 'siso f32 samples')
    nbad = assert_decisions(_t(hat), ref_hat, om, ref_r, exact=False, what='siso f32')
    ref_cnt = OL.counters(idx, ref_hat, 6)
    assert abs(int(cnt[0]) - int(ref_cnt[0])) <= nbad and cnt[2] == n and cnt[3] == 6 * n


def test_c2_golden(golden):
    from pyphysim_b200 import links
    g = golden('links')
    pm = product_modem('qam', 64)
    idx, h, nz = OL.draws_siso_flat(SEED, np.arange(4096), 6)
    assert np.array_equal(idx, g['c2_idx'])
    cnt, hat = links.link_siso_flat(pm, 1 / md.dB2Linear(10.0), 4096, dtype='f64',
                                    draws=(cuda(idx.astype(np.uint8)), cuda(h), cuda(nz)), want_idx=True)
    assert np.array_equal(_t(hat), g['c2_hat'])              # == the reference's own output


def test_siso_flat_empty_and_ragged():
    from pyphysim_b200 import links
    pm = product_modem('qam', 16)
    assert np.array_equal(links.link_siso_flat(pm, 0.1, 0), [0, 0, 0, 0])
    for n in (1, 2, 3, 5, 255, 257):
        c = links.link_siso_flat(pm, 0.0, n, dtype='f32')
        assert list(c) == [0, 0, n, 4 * n]                   # no noise -> no errors
    c = links.link_siso_flat(pm, 0.5, 10 ** 6, dtype='f32')
    assert 0.2 < c[0] / c[2] < 0.9 and c[1] >= c[0]


def test_host_entry_point_matches_device_path():
    from pyphysim_b200 import links
    pm = product_modem('qam', 64)
    n, nv = 300001, 0.05
    d = links.draw_siso_flat(pm, n, seed=SEED, dtype='f32')
    c_dev, hat_dev = links.link_siso_flat(pm, nv, n, dtype='f32', draws=d, want_idx=True)
    host = tuple(t.cpu().pin_memory() for t in d)
    c_host, hat_host = links.link_siso_flat_host(pm, nv, n, dtype='f32', draws=host, want_idx=True)
    assert np.array_equal(c_dev, c_host) and np.array_equal(_t(hat_dev), hat_host.numpy())
    c_f = links.link_siso_flat_host(pm, nv, n, dtype='f32', seed=SEED)
    assert np.array_equal(c_f, c_dev)


# ------------------------------------------------------------------ C4: Alamouti
@pytest.mark.parametrize('kind,M', [('qpsk', 4), ('qam', 16)])
def test_alamouti_c4_f32_kernel(kind, M):
    """The compile-time-shape float32 kernel of the C4 shape (Nr = 2, one codeword): stream mode == fused mode on
    the device's own draws (one arithmetic for both), == the host-buffer entry point in both modes, samples within
    1e-5 of the oracle, mismatches only on decision boundaries; misaligned tensors are refused, not faulted on."""
    from pyphysim_b200 import links
    import torch
    pm, om = product_modem(kind, M), oracle_modem(kind, M)
    n, nv = 50001, 1 / md.dB2Linear(9.0)
    d = links.draw_flat_mimo(pm, n, Nr=2, Nt=2, num_symbols=2, n_data=2, seed=SEED, first_unit=77, dtype='f32')
    c_s, hat_s, dec_s = links.link_alamouti(pm, nv, n, draws=d, dtype='f32', want_idx=True, want_samples=True)
    c_f, hat_f = links.link_alamouti(pm, nv, n, seed=SEED, first_unit=77, dtype='f32', want_idx=True)
    assert np.array_equal(c_s, c_f) and torch.equal(hat_s, hat_f)
    c_n = links.link_alamouti(pm, nv, n, draws=d, dtype='f32')              # no outputs: QPSK skips the gain
    assert np.array_equal(c_n, c_s)
    ref_hat, ref_dec = OL.alamouti(om, _t(d[0]).astype(np.int64), _t(d[1]).astype(complex), _t(d[2]).astype(complex), nv)
    assert_samples_close(_t(dec_s), ref_dec, 1e-5, 'alamouti22 f32')
    assert_decisions(_t(hat_s), ref_hat, om, ref_dec, exact=False, eps=2e-5, what='alamouti22 f32')
    host = tuple(t.cpu().pin_memory() for t in d)
    c_h, hat_h = links.link_alamouti_host(pm, nv, n, draws=host, dtype='f32', want_idx=True)
    assert np.array_equal(c_h, c_s) and np.array_equal(hat_h.numpy(), _t(hat_s))
    assert np.array_equal(links.link_alamouti_host(pm, nv, n, seed=SEED, first_unit=77, dtype='f32'), c_s)
    with pytest.raises(ValueError, match='aligned'):
        links.link_alamouti(pm, nv, 8, draws=(d[0][:8], d[1].view(torch.float32).reshape(-1)[2:2 + 64].view(torch.complex64),
                                               d[2][:8]), dtype='f32')


def test_blast_and_precoded_host_entry_points():
    from pyphysim_b200 import links
    pm = product_modem('qam', 16)
    n, nv = 20001, 0.02
    d = links.draw_flat_mimo(pm, n, Nr=2, Nt=2, num_symbols=3, n_data=6, seed=SEED, dtype='f32')
    c_dev, hat_dev = links.link_blast(pm, nv, n, Nr=2, Nt=2, num_symbols=3, filter_noise_var=nv, draws=d, dtype='f32',
                                      want_idx=True)
    host = tuple(t.cpu().pin_memory() for t in d)
    c_h, hat_h = links.link_blast_host(pm, nv, n, Nr=2, Nt=2, num_symbols=3, filter_noise_var=nv, draws=host,
                                       dtype='f32', want_idx=True)
    assert np.array_equal(c_h, c_dev) and np.array_equal(hat_h.numpy(), _t(hat_dev))
    assert np.array_equal(links.link_blast_host(pm, nv, n, Nr=2, Nt=2, num_symbols=3, filter_noise_var=nv, seed=SEED,
                                                dtype='f32'), c_dev)
    d = links.draw_flat_mimo(pm, n, Nr=4, Nt=4, num_symbols=2, n_data=8, seed=SEED, dtype='f32')
    for scheme in ('svd', 'gmd'):
        c_dev = links.link_precoded(pm, nv, n, scheme=scheme, Nr=4, Nt=4, num_symbols=2, draws=d, dtype='f32')
        c_h = links.link_precoded_host(pm, nv, n, scheme=scheme, Nr=4, Nt=4, num_symbols=2,
                                       draws=tuple(t.cpu() for t in d), dtype='f32')
        assert np.array_equal(c_h, c_dev)
        assert np.array_equal(links.link_precoded_host(pm, nv, n, scheme=scheme, Nr=4, Nt=4, num_symbols=2, seed=SEED,
                                                       dtype='f32'), c_dev)


@pytest.mark.parametrize('Nr,S,kind,M', [(2, 2, 'qpsk', 4), (1, 4, 'qam', 16), (3, 6, 'psk', 8)])
def test_alamouti_f64_bit_exact_vs_oracle(Nr, S, kind, M):
    from pyphysim_b200 import links
    import torch
    pm, om = product_modem(kind, M), oracle_modem(kind, M)
    n, nv = 3001, 1 / md.dB2Linear(8.0)
    idx, H, nz = OL.draws_flat_mimo(SEED, np.arange(5, 5 + n), om.bits, Nr, 2, S, S)
    ref_hat, ref_dec = OL.alamouti(om, idx, H, nz, nv)
    cnt, hat, dec = links.link_alamouti(pm, nv, n, Nr=Nr, num_symbols=S, dtype='f64',
                                        draws=(cuda(idx.astype(np.uint8)), cuda(H), cuda(nz)),
                                        want_idx=True, want_samples=True)
    assert_samples_close(_t(dec), ref_dec, 1e-11, 'alamouti f64')
    assert_decisions(_t(hat), ref_hat, om, ref_dec, exact=True, what='alamouti f64')
    assert np.array_equal(cnt, OL.counters(idx, ref_hat, om.bits))
    d = links.draw_flat_mimo(pm, n, Nr=Nr, Nt=2, num_symbols=S, n_data=S, seed=SEED, first_unit=5, dtype='f64')
    c_s, hat_s = links.link_alamouti(pm, nv, n, Nr=Nr, num_symbols=S, dtype='f64', draws=d, want_idx=True)
    c_f, hat_f = links.link_alamouti(pm, nv, n, Nr=Nr, num_symbols=S, dtype='f64', seed=SEED, first_unit=5,
                                     want_idx=True)
    assert np.array_equal(c_s, c_f) and torch.equal(hat_s, hat_f)


def test_c4_golden_and_f32(golden):
    from pyphysim_b200 import links
    g = golden('links')
    pm, om = product_modem('qpsk'), oracle_modem('qpsk')
    idx, H, nz = OL.draws_flat_mimo(SEED, np.arange(64), 2, 2, 2, 2, 2)
    nv = 1 / md.dB2Linear(10.0)
    cnt, hat = links.link_alamouti(pm, nv, 64, dtype='f64',
                                   draws=(cuda(idx.astype(np.uint8)), cuda(H), cuda(nz)), want_idx=True)
    assert np.array_equal(_t(hat), g['c4_hat'])
    n = 100000
    d = links.draw_flat_mimo(pm, n, Nr=2, Nt=2, num_symbols=2, n_data=2, seed=SEED, dtype='f32')
    cnt, hat, dec = links.link_alamouti(pm, nv, n, dtype='f32', seed=SEED, want_idx=True, want_samples=True)
    ref_hat, ref_dec = OL.alamouti(om, _t(d[0]).astype(np.int64), _t(d[1]).astype(complex),
                                   _t(d[2]).astype(complex), nv)
    assert_samples_close(_t(dec), ref_dec, 1e-5, 'alamouti f32')
    assert_decisions(_t(hat), ref_hat, om, ref_dec, exact=False, what='alamouti f32')


# ------------------------------------------------------------------ Blast ZF / MMSE
@pytest.mark.parametrize('Nr,Nt,S,fnv', [(2, 2, 1, 0.0), (4, 3, 3, 0.0), (4, 4, 2, 0.02), (2, 1, 4, 0.1),
                                         (3, 4, 2, 0.05)])
def test_blast_f64_vs_oracle(Nr, Nt, S, fnv):
    from pyphysim_b200 import links
    import torch
    pm, om = product_modem('qam', 16), oracle_modem('qam', 16)
    n, nv = 2000, 0.02
    idx, H, nz = OL.draws_flat_mimo(SEED, np.arange(9, 9 + n), 4, Nr, Nt, S, S * Nt)
    ref_hat, ref_dec = OL.blast_flat(om, idx, H, nz, nv, fnv)
    cnt, hat, dec = links.link_blast(pm, nv, n, Nr=Nr, Nt=Nt, num_symbols=S, filter_noise_var=fnv,
                                     dtype='f64', draws=(cuda(idx.astype(np.uint8)), cuda(H), cuda(nz)),
                                     want_idx=True, want_samples=True)
    # ZF of an ill-conditioned H amplifies rounding: compare relative to the amplified scale
    assert_samples_close(_t(dec), ref_dec, 1e-7, 'blast f64')
    assert_decisions(_t(hat), ref_hat, om, ref_dec, exact=False, eps=1e-6, what='blast f64')
    d = links.draw_flat_mimo(pm, n, Nr=Nr, Nt=Nt, num_symbols=S, n_data=S * Nt, seed=SEED, first_unit=9,
                             dtype='f64')
    c_s, hat_s = links.link_blast(pm, nv, n, Nr=Nr, Nt=Nt, num_symbols=S, filter_noise_var=fnv, dtype='f64',
                                  draws=d, want_idx=True)
    c_f, hat_f = links.link_blast(pm, nv, n, Nr=Nr, Nt=Nt, num_symbols=S, filter_noise_var=fnv, dtype='f64',
                                  seed=SEED, first_unit=9, want_idx=True)
    assert np.array_equal(c_s, c_f) and torch.equal(hat_s, hat_f)


def test_blast_errors():
    from pyphysim_b200 import links
    pm = product_modem('qam', 16)
    with pytest.raises(NotImplementedError):
        links.link_blast(pm, 0.1, 10, Nr=2, Nt=3, filter_noise_var=0.0)      # ZF needs Nt <= Nr
    with pytest.raises(ValueError):
        links.link_blast(pm, 0.1, 10, Nr=2, Nt=2, filter_noise_var=-1.0)
    with pytest.raises(ValueError):
        links.link_alamouti(pm, 0.1, 10, num_symbols=3)
