"""Oracle vs fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import math

import numpy as np

from oracle import fading, links, mimo, modulators as md, ofdm, philox

SEED = 0xC0FFEE
TOL = dict(rtol=1e-11, atol=1e-12)


def test_constellations(golden):
    g = golden('constellations')
    for M in (4, 16, 64, 256):
        np.testing.assert_allclose(md.qam_constellation(M), g['qam%d' % M], **TOL)
    for M in (2, 4, 8, 16):
        np.testing.assert_allclose(md.psk_constellation(M), g['psk%d' % M], **TOL)
    np.testing.assert_allclose(md.psk_constellation(8, 0.3), g['psk8_off'], **TOL)
    np.testing.assert_allclose(md.psk_raw(8, 0.2), g['psk8_setoffset'], **TOL)
    np.testing.assert_allclose(md.qpsk_constellation(), g['qpsk'], **TOL)
    assert np.array_equal(md.bpsk_constellation(), g['bpsk'])
    assert np.array_equal(md.binary2gray(np.arange(16)), g['gray16'])
    assert np.array_equal(md.gray2binary(np.arange(16)), g['ungray16'])


def test_demap_and_bit_errors(golden):
    g = golden('demap')
    mods = {'qam16': links.Modem('qam', 16), 'qam64': links.Modem('qam', 64),
            'qam256': links.Modem('qam', 256), 'psk8': links.Modem('psk', 8),
            'qpsk': links.Modem('psk', 4, np.pi / 4), 'bpsk': links.Modem('bpsk')}
    for name, m in mods.items():
        hat = m.demodulate(g[name + '_r'])
        assert np.array_equal(hat, g[name + '_hat']), name
        assert md.count_bit_errors(g[name + '_idx'], hat) == int(g[name + '_biterr'])
        if name.startswith('qam') or name == 'psk8':
            assert int(g[name + '_biterr']) > 0      # the fixture exercises errors
    assert np.array_equal(md.count_bits(g['count_bits_in']), g['count_bits_out'])


def test_ofdm(golden):
    g = golden('ofdm')
    for tag in 'abcd':
        f, c, u = (int(v) for v in g[tag + '_params'])
        assert np.array_equal(ofdm.used_subcarrier_indexes(f, u), g[tag + '_bins'])
        np.testing.assert_allclose(ofdm.modulate(g[tag + '_x'], f, c, u), g[tag + '_mod'], **TOL)
        np.testing.assert_allclose(ofdm.demodulate(g[tag + '_r'], f, c, u), g[tag + '_demod'],
                                   **TOL)


def test_profiles_and_jakes(golden):
    g = golden('fading')
    profs = {'tu': fading.COST259_TU, 'ra': fading.COST259_RA, 'ht': fading.COST259_HT}
    for pn, prof in profs.items():
        for tn, Ts in (('2048', 1 / (15e3 * 2048)), ('1024', 1 / (15e3 * 1024)),
                       ('128', 1 / (15e3 * 128))):
            p, d = fading.discretize_profile(prof[0], prof[1], Ts)
            assert np.array_equal(d, g['%s_%s_delays' % (pn, tn)])
            np.testing.assert_allclose(p, g['%s_%s_powers' % (pn, tn)], rtol=1e-14)
    # constructor emitted one sample at t=0, so generation resumes at t0 = Ts
    h1, t1 = fading.jakes_samples(g['j0_phi'], g['j0_psi'], 100.0, 1e-3, 1e-3, 50)
    np.testing.assert_allclose(h1, g['j0_h1'][0] if g['j0_h1'].ndim > 1 else g['j0_h1'], **TOL)
    h2, t2 = fading.jakes_samples(g['j0_phi'], g['j0_psi'], 100.0, 1e-3, t1 + 7 * 1e-3, 20)
    np.testing.assert_allclose(h2, g['j0_h2'][0] if g['j0_h2'].ndim > 1 else g['j0_h2'], **TOL)
    assert abs(t2 - float(g['j0_t_end'])) < 1e-12
    h, _ = fading.jakes_samples(g['j1_phi'], g['j1_psi'], 30.0, 5e-6, 5e-6, 40)
    np.testing.assert_allclose(h, g['j1_h'], **TOL)


def test_phase_layout_matches_fixture(golden):
    g = golden('tdl')
    phi, psi = philox.jakes_phases(SEED, [400], (20, g['s_delays'].size))
    assert np.array_equal(phi[0], g['s_phi']) and np.array_equal(psi[0], g['s_psi'])
    phi, psi = philox.jakes_phases(SEED, [401], g['m_phi'].shape)
    assert np.array_equal(phi[0], g['m_phi']) and np.array_equal(psi[0], g['m_psi'])


def test_tdl_siso_chain(golden):
    g = golden('tdl')
    fft, cp, used, nsym = (int(v) for v in g['s_params'])
    Ts, Fd = float(g['s_Ts']), float(g['s_Fd'])
    p, d = fading.discretize_profile(*fading.COST259_TU, Ts)
    assert np.array_equal(d, g['s_delays'])
    np.testing.assert_allclose(p, g['s_powers'], rtol=1e-14)
    q = md.qam_constellation(64)
    tx = ofdm.modulate(md.modulate(q, g['s_idx']), fft, cp, used)
    np.testing.assert_allclose(tx, g['s_tx'], **TOL)
    h, _ = fading.jakes_samples(g['s_phi'], g['s_psi'], Fd, Ts, Ts, tx.size)
    taps = fading.tdl_taps(h, p)
    np.testing.assert_allclose(taps, g['s_taps'], **TOL)
    rx = fading.tdl_corrupt(tx, taps, d)
    np.testing.assert_allclose(rx, g['s_rx'], **TOL)
    Y = ofdm.demodulate((rx + math.sqrt(float(g['s_nv'])) * g['s_noise'])[:tx.size], fft, cp, used)
    np.testing.assert_allclose(Y, g['s_Y'], **TOL)
    # the reference's per-sample-FFT equaliser and the FFT(mean taps) restatement
    H_ref = ofdm.mean_freq_response_reference(taps, d, fft, nsym)
    H_fast = ofdm.mean_freq_response(taps, d, fft, nsym)
    np.testing.assert_allclose(H_fast, H_ref, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(ofdm.onetap_equalize(Y, H_ref, fft, used), g['s_eq'], **TOL)
    np.testing.assert_allclose(ofdm.onetap_equalize(Y, H_fast, fft, used), g['s_eq'],
                               rtol=1e-10, atol=1e-11)
    assert np.array_equal(md.demodulate(q, g['s_eq']), g['s_hat'])


def test_tdl_mimo_corrupt(golden):
    g = golden('tdl')
    Ts, Fd = float(g['m_Ts']), float(g['m_Fd'])
    p, d = fading.discretize_profile(*fading.COST259_RA, Ts)
    assert np.array_equal(d, g['m_delays'])
    h, _ = fading.jakes_samples(g['m_phi'], g['m_psi'], Fd, Ts, Ts, g['m_x'].shape[1])
    taps = fading.tdl_taps(h, p)
    np.testing.assert_allclose(taps, g['m_taps'], **TOL)
    np.testing.assert_allclose(fading.tdl_corrupt(g['m_x'], taps, d), g['m_y'], **TOL)
    np.testing.assert_allclose(np.fft.fft(ofdm.dense_taps(taps, d), 64, axis=0)[..., ::50],
                               g['m_freq'], **TOL)


def test_mimo(golden):
    g = golden('mimo')
    for name in ('h43', 'h44', 'h22'):
        H = g[name]
        np.testing.assert_allclose(mimo.blast_encode(g[name + '_s'], H.shape[1]), g[name + '_x'],
                                   **TOL)
        np.testing.assert_allclose(mimo.blast_decode(g[name + '_y'], H), g[name + '_zf'],
                                   rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(mimo.blast_decode(g[name + '_y'], H, 0.01), g[name + '_mmse'],
                                   rtol=1e-9, atol=1e-10)
    for name in ('a22', 'a32', 'a12'):
        np.testing.assert_allclose(mimo.alamouti_encode(g[name + '_s']), g[name + '_x'], **TOL)
        np.testing.assert_allclose(mimo.alamouti_decode(g[name + '_y'], g[name]),
                                   g[name + '_dec'], **TOL)


def test_mimo_schemes(golden):
    """next-3: MRT / MRC / SVDMimo / GMDMimo, gmd() and the post-processing SINRs against the reference
    (tests/golden/make_golden_mimo_schemes.py).  SVD-based outputs use numpy's SVD exactly as the reference
    does, so they are compared tightly; the gauge-fixed SVD the CUDA path uses is checked against them
    through the gauge-invariant products."""
    g = golden('mimo_schemes')
    nv = float(g['noise_var'])
    for n in (2, 3, 4):
        pre = 'sq%d_' % n
        H, x, noise = g[pre + 'H'], g[pre + 'x'], g[pre + 'noise']
        Q, R, P = mimo.gmd(g[pre + 'U'], g[pre + 'S'], g[pre + 'Vh'])
        for got, ref in ((Q, g[pre + 'Q']), (R, g[pre + 'R']), (P, g[pre + 'P'])):
            assert np.array_equal(got, ref)                       # same arithmetic on the same SVD: bit exact
        np.testing.assert_allclose(Q.dot(R).dot(P.conj().T), H, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(np.diag(R), np.prod(g[pre + 'S']) ** (1.0 / n) * np.ones(n), rtol=1e-13)
        for name, Wf, Gf in (('svd', mimo.svd_precoder, lambda h: mimo.svd_receive_filter(h)),
                             ('gmd', mimo.gmd_precoder, lambda h: mimo.gmd_receive_filter(h, nv))):
            W, G = Wf(H), Gf(H)
            np.testing.assert_allclose(W, g[pre + name + '_W'], **TOL)
            np.testing.assert_allclose(G, g[pre + name + '_G'], rtol=1e-9, atol=1e-10)
            enc = mimo.precoded_encode(x, W)
            np.testing.assert_allclose(enc, g[pre + name + '_enc'], **TOL)
            np.testing.assert_allclose(G.dot(H.dot(enc) + noise).reshape(-1), g[pre + name + '_dec'], rtol=1e-9, atol=1e-10)
            sin = mimo.post_processing_linear_sinrs(H, W, G, nv)
            # the reference's calc_linear_SINRs returns dB (it calls calc_post_processing_SINRs, mimo.py:311-333)
            np.testing.assert_allclose(10 * np.log10(sin), g[pre + name + '_sinr_lin'], rtol=1e-8, atol=1e-9)
            assert int(g[pre + name + '_layers']) == n
        # gauge-fixed SVD: same singular values, and the SVD filters differ from numpy's by one unit phase
        # per stream that cancels in G H W
        U, S, V = mimo.svd_canonical(H)
        np.testing.assert_allclose(S, g[pre + 'S'], rtol=1e-13)
        np.testing.assert_allclose(U.dot(np.diag(S)).dot(V.conj().T), H, rtol=1e-12, atol=1e-13)
        Wc, Gc = V / math.sqrt(n), np.diag(1 / S).dot(U.conj().T) * math.sqrt(n)
        np.testing.assert_allclose(Gc.dot(H).dot(Wc), np.eye(n), rtol=1e-10, atol=1e-11)
        D = np.diag(Wc.conj().T.dot(g[pre + 'svd_W'])) * n
        np.testing.assert_allclose(np.abs(D), np.ones(n), rtol=1e-12)
        np.testing.assert_allclose(mimo.post_processing_linear_sinrs(H, Wc, Gc, nv),
                                   mimo.post_processing_linear_sinrs(H, g[pre + 'svd_W'], g[pre + 'svd_G'], nv), rtol=1e-9)
    Q, R, P = mimo.gmd(g['tall_U'], g['tall_S'], g['tall_Vh'])
    assert np.array_equal(Q, g['tall_Q']) and np.array_equal(R, g['tall_R']) and np.array_equal(P, g['tall_P'])
    Q, R, P = mimo.gmd(g['tol_U'], g['tol_S'], g['tol_Vh'], float(g['tol_tol']))
    assert np.array_equal(Q, g['tol_Q']) and np.array_equal(R, g['tol_R']) and np.array_equal(P, g['tol_P'])
    for nt in (2, 3, 4):
        pre = 'mrt%d_' % nt
        h, x, noise = g[pre + 'h'], g[pre + 'x'], g[pre + 'noise']
        np.testing.assert_allclose(mimo.mrt_precoder(h), g[pre + 'W'], **TOL)
        np.testing.assert_allclose(mimo.mrt_receive_filter(h), g[pre + 'G'], **TOL)
        enc = mimo.mrt_encode(x, h)
        np.testing.assert_allclose(enc, g[pre + 'enc'], **TOL)
        np.testing.assert_allclose(mimo.mrt_decode(h.dot(enc) + noise, h), g[pre + 'dec'], **TOL)
        sin = mimo.post_processing_linear_sinrs(h, mimo.mrt_precoder(h), mimo.mrt_receive_filter(h), nv)
        np.testing.assert_allclose(10 * np.log10(sin), g[pre + 'sinr_lin'].reshape(-1), rtol=1e-9)
    # MRC = Blast with a column channel (mimo.py:786-826)
    h = g['mrc_h'][:, None]
    enc = mimo.blast_encode(g['mrc_x'], 1)
    np.testing.assert_allclose(enc, g['mrc_enc'], **TOL)
    np.testing.assert_allclose(mimo.blast_decode(h.dot(enc) + g['mrc_noise'], h, nv), g['mrc_dec'], rtol=1e-9, atol=1e-10)
    assert int(g['mrc_layers']) == 1
    H = g['mrc2_H']
    enc = mimo.blast_encode(g['mrc2_x'], 3)
    np.testing.assert_allclose(mimo.blast_decode(H.dot(enc) + g['mrc2_noise'], H), g['mrc2_dec'], rtol=1e-9, atol=1e-10)


def test_link_c3_full_size(golden):
    g = golden('links')
    m = links.Modem('qam', 64)
    cfg = links.OfdmTdlConfig(m, 1024, 72, 1024, noise_var=1 / md.dB2Linear(20.0))
    idx, phi, psi, noise = links.draws_ofdm_tdl(cfg, SEED, g['c3_units'])
    assert np.array_equal(idx, g['c3_idx'])
    for u in range(2):
        for ref_eq in (True, False):
            hat, det = links.ofdm_tdl_frame(cfg, idx[u], phi[u], psi[u], noise[u],
                                            reference_equalizer=ref_eq, detail=True)
            np.testing.assert_allclose(det['eq'], g['c3_eq'][u], rtol=1e-9, atol=1e-10)
            assert np.array_equal(hat, g['c3_hat'][u])
    c = links.counters(idx, g['c3_hat'], 6)
    assert 0 < c[0] < c[2]                           # the link does make errors


def test_link_mimo2x2_ofdm(golden):
    g = golden('links')
    m = links.Modem('qam', 16)
    cfg = links.OfdmTdlConfig(m, 256, 18, 200, n_sym=2, Nr=2, Nt=2, Fd=300.0,
                              noise_var=1 / md.dB2Linear(25.0))
    idx, phi, psi, noise = links.draws_ofdm_tdl(cfg, SEED, g['m2_units'])
    assert np.array_equal(idx, g['m2_idx'])
    for u in range(2):
        hat, det = links.ofdm_tdl_frame(cfg, idx[u], phi[u], psi[u], noise[u], detail=True)
        np.testing.assert_allclose(det['eq'], g['m2_eq'][u], rtol=1e-8, atol=1e-9)
        assert np.array_equal(hat, g['m2_hat'][u])


def test_links_flat(golden):
    g = golden('links')
    idx, H, n = links.draws_flat_mimo(SEED, np.arange(64), 2, 2, 2, 2, 2)
    assert np.array_equal(idx, g['c4_idx'])
    hat, _ = links.alamouti(links.Modem('psk', 4, np.pi / 4), idx, H, n, 1 / md.dB2Linear(10.0))
    assert np.array_equal(hat, g['c4_hat'])
    idx, h, n = links.draws_siso_flat(SEED, np.arange(4096), 6)
    assert np.array_equal(idx, g['c2_idx'])
    hat, _ = links.siso_flat(links.Modem('qam', 64), idx, h, n, 1 / md.dB2Linear(10.0))
    assert np.array_equal(hat, g['c2_hat'])
    assert 0 < np.sum(hat != idx) < idx.size
