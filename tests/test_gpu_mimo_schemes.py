"""GPU parity of SURVEY.md §8f row next-3: the batched small SVD / GMD, MRT / MRC / SVDMimo / GMDMimo with
the reference's API, and the fused precoded link — against the oracle (gauge-fixed SVD) and against the
fixture produced by the unmodified reference (tests/golden/make_golden_mimo_schemes.py).  Everything goes
through the C ABI."""
import math

import numpy as np
import pytest

from oracle import links as OL
from oracle import mimo as OM
from oracle import philox

from gpu_util import assert_decisions, assert_samples_close, cuda, oracle_modem, product_modem

pytestmark = pytest.mark.gpu
SEED = 0xC0FFEE


def _t(x):
    return x.cpu().numpy()


def _channels(first, n, Nr, Nt):
    return philox.cnormal(SEED, 1, np.arange(first, first + n), Nr * Nt).reshape(n, Nr, Nt)


# ------------------------------------------------------------------ decompositions
@pytest.mark.parametrize('Nr,Nt', [(2, 2), (3, 3), (4, 4), (4, 2), (3, 1), (1, 1), (4, 3)])
def test_svd_batch_vs_numpy(Nr, Nt):
    import torch
    from pyphysim_b200 import _lib
    lib = _lib.load()
    n = 3000
    H = _channels(40, n, Nr, Nt)
    h = cuda(H)
    U = torch.empty((n, Nr, Nt), dtype=torch.complex128, device='cuda')
    S = torch.empty((n, Nt), dtype=torch.float64, device='cuda')
    V = torch.empty((n, Nt, Nt), dtype=torch.complex128, device='cuda')
    _lib.check(lib.b200phy_svd(_lib.ptr(h), n, Nr, Nt, _lib.ptr(U), _lib.ptr(S), _lib.ptr(V), None))
    U, S, V = _t(U), _t(S), _t(V)
    np.testing.assert_allclose(S, np.linalg.svd(H, compute_uv=False), rtol=1e-12, atol=1e-13)
    rec = np.einsum('nik,nk,njk->nij', U, S, V.conj())
    np.testing.assert_allclose(rec, H, rtol=0, atol=1e-12)
    eye = np.broadcast_to(np.eye(Nt), (n, Nt, Nt))
    np.testing.assert_allclose(np.einsum('nki,nkj->nij', U.conj(), U), eye, atol=1e-12)
    np.testing.assert_allclose(np.einsum('nki,nkj->nij', V.conj(), V), eye, atol=1e-12)
    # the fixed gauge: every pair equals the oracle's gauge-fixed numpy SVD (well separated singular values)
    gap_ok = 0
    for u in range(200):
        Uc, Sc, Vc = OM.svd_canonical(H[u])
        gaps = np.abs(np.diff(Sc)) if Nt > 1 else np.array([1.0])
        if gaps.min() < 1e-3:
            continue
        gap_ok += 1
        np.testing.assert_allclose(V[u], Vc, atol=1e-9)
        np.testing.assert_allclose(U[u], Uc, atol=1e-9)
    assert gap_ok > 150


@pytest.mark.parametrize('Nr,Nt', [(2, 2), (3, 3), (4, 4), (4, 2), (4, 3)])
def test_gmd_batch_vs_oracle(Nr, Nt):
    """Same U, S, V in, same Q, R, P out as the restatement of util.misc.gmd (itself bit-exact against the
    reference, tests/test_oracle_golden.py)."""
    import torch
    from pyphysim_b200 import _lib
    lib = _lib.load()
    n = 500
    H = _channels(77, n, Nr, Nt)
    Us, Ss, Vs = zip(*(OM.svd_canonical(h) for h in H))
    U, S, V = np.stack(Us), np.stack(Ss), np.stack(Vs)
    Q = torch.empty((n, Nr, Nt), dtype=torch.complex128, device='cuda')
    R = torch.empty((n, Nt, Nt), dtype=torch.float64, device='cuda')
    P = torch.empty((n, Nt, Nt), dtype=torch.complex128, device='cuda')
    u_d, s_d, v_d = cuda(U), cuda(S), cuda(V)                      # keep the inputs alive across the call
    _lib.check(lib.b200phy_gmd(_lib.ptr(u_d), _lib.ptr(s_d), _lib.ptr(v_d), n, Nr, Nt, _lib.ptr(Q),
                               _lib.ptr(R), _lib.ptr(P), None))
    Q, R, P = _t(Q), _t(R), _t(P)
    for u in range(n):
        q, r, p = OM.gmd(U[u], S[u], V[u].conj().T)
        np.testing.assert_allclose(R[u], r[:Nt], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(Q[u], q, atol=1e-10)
        np.testing.assert_allclose(P[u], p, atol=1e-10)
        np.testing.assert_allclose(Q[u].dot(R[u]).dot(P[u].conj().T), H[u], atol=1e-11)
        np.testing.assert_allclose(np.diag(R[u]), np.prod(S[u]) ** (1.0 / Nt) * np.ones(Nt), rtol=1e-12)
        assert np.all(np.tril(R[u], -1) == 0)


def test_util_gmd_matches_reference_fixture(golden):
    from pyphysim_b200.util import misc
    g = golden('mimo_schemes')
    for pre in ('sq2_', 'sq3_', 'sq4_', 'tall_'):
        Q, R, P = misc.gmd(g[pre + 'U'], g[pre + 'S'], g[pre + 'Vh'])
        np.testing.assert_allclose(Q, g[pre + 'Q'], atol=1e-12)
        np.testing.assert_allclose(R, g[pre + 'R'], atol=1e-12)
        np.testing.assert_allclose(P, g[pre + 'P'], atol=1e-12)
    with pytest.raises(NotImplementedError):
        misc.gmd(g['tol_U'], g['tol_S'], g['tol_Vh'], float(g['tol_tol']))


# ------------------------------------------------------------------ facade classes vs the reference fixture
def test_mrt_mrc_classes(golden):
    from pyphysim_b200 import mimo
    g = golden('mimo_schemes')
    nv = float(g['noise_var'])
    for k, nt in enumerate((2, 3, 4)):
        pre = 'mrt%d_' % nt
        h = g[pre + 'h']
        obj = mimo.MRT(h[0] if k == 0 else h)
        assert (obj.Nr, obj.Nt, obj.getNumberOfLayers()) == (1, nt, 1)
        enc = obj.encode(g[pre + 'x'])
        np.testing.assert_allclose(enc, g[pre + 'enc'], rtol=1e-12, atol=1e-13)
        dec = obj.decode(h.dot(enc) + g[pre + 'noise'])
        assert dec.ndim == 1
        np.testing.assert_allclose(dec, g[pre + 'dec'], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(obj.calc_linear_SINRs(nv), g[pre + 'sinr_lin'].reshape(-1), rtol=1e-10)
    with pytest.raises(ValueError):
        mimo.MRT(np.ones((2, 3), dtype=complex))
    obj = mimo.MRC(g['mrc_h'])
    obj.set_noise_var(nv)
    assert (obj.Nr, obj.Nt, obj.getNumberOfLayers()) == (4, 1, 1)
    enc = obj.encode(g['mrc_x'])
    np.testing.assert_allclose(enc, g['mrc_enc'], rtol=1e-12)
    np.testing.assert_allclose(obj.decode(g['mrc_h'][:, None].dot(enc) + g['mrc_noise']), g['mrc_dec'], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(obj.calc_linear_SINRs(nv), g['mrc_sinr_lin'], rtol=1e-9)
    obj = mimo.MRC(g['mrc2_H'])
    obj.set_noise_var(None)
    enc = obj.encode(g['mrc2_x'])
    np.testing.assert_allclose(obj.decode(g['mrc2_H'].dot(enc) + g['mrc2_noise']), g['mrc2_dec'], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize('n', [2, 3, 4])
def test_svd_gmd_classes(golden, n):
    from pyphysim_b200 import mimo
    g = golden('mimo_schemes')
    nv = float(g['noise_var'])
    pre = 'sq%d_' % n
    H, x, noise = g[pre + 'H'], g[pre + 'x'], g[pre + 'noise']
    # ---- SVDMimo: equal to the reference up to one unit phase per stream (the SVD gauge)
    obj = mimo.SVDMimo(H)
    assert obj.getNumberOfLayers() == n
    W, G = obj._calc_precoder(H), obj._calc_receive_filter(H)
    D = np.diag(g[pre + 'svd_W'].conj().T.dot(W)) * n                 # W = W_ref diag(D)
    np.testing.assert_allclose(np.abs(D), np.ones(n), rtol=1e-11)
    np.testing.assert_allclose(W, g[pre + 'svd_W'] * D[np.newaxis, :], atol=1e-11)
    np.testing.assert_allclose(G, D.conj()[:, np.newaxis] * g[pre + 'svd_G'], rtol=1e-9, atol=1e-10)
    enc = obj.encode(x)
    np.testing.assert_allclose(enc, W.dot(x.reshape(n, -1)), atol=1e-12)
    dec = obj.decode(H.dot(enc) + noise)
    ref = x.reshape(n, -1) + D.conj()[:, np.newaxis] * (g[pre + 'svd_dec'].reshape(n, -1) - x.reshape(n, -1))
    np.testing.assert_allclose(dec, ref.reshape(-1), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(obj.calc_linear_SINRs(nv), g[pre + 'svd_sinr_lin'], rtol=1e-8)      # gauge invariant
    np.testing.assert_allclose(obj.calc_SINRs(nv), g[pre + 'svd_sinr_dB'], rtol=1e-8)
    with pytest.raises(ValueError):
        obj.encode(np.ones(5 * n + 1, dtype=complex))
    with pytest.raises(ValueError):
        mimo.SVDMimo(np.ones((4, 2), dtype=complex))
    # ---- GMDMimo: the decomposition depends on the SVD gauge; compare with the oracle in the same gauge and
    # with the reference through what is gauge invariant (R, reconstruction, SINRs, noiseless round trip)
    obj = mimo.GMDMimo(H)
    obj.set_noise_var(nv)
    Uc, Sc, Vc = OM.svd_canonical(H)
    Q, R, P = OM.gmd(Uc, Sc, Vc.conj().T)
    W, G = obj._calc_precoder(H), obj._calc_receive_filter(H, nv)
    np.testing.assert_allclose(W, P / math.sqrt(n), atol=1e-10)
    np.testing.assert_allclose(G, OM.blast_receive_filter(Q.dot(R), nv), rtol=1e-8, atol=1e-9)
    enc = obj.encode(x)
    dec = obj.decode(H.dot(enc) + noise)
    np.testing.assert_allclose(dec, G.dot(H.dot(W.dot(x.reshape(n, -1))) + noise).reshape(-1), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(np.diag(R), np.diag(g[pre + 'R']), rtol=1e-12)
    np.testing.assert_allclose(obj.calc_linear_SINRs(nv), g[pre + 'gmd_sinr_lin'], rtol=1e-7)
    obj.set_noise_var(None)
    np.testing.assert_allclose(obj.decode(H.dot(obj.encode(x))), x, atol=1e-9)        # zero-forcing round trip
    # post-processing SINRs as free functions
    np.testing.assert_allclose(mimo.calc_post_processing_SINRs(H, g[pre + 'svd_W'], g[pre + 'svd_G'], nv),
                               g[pre + 'svd_sinr_lin'], rtol=1e-10)


# ------------------------------------------------------------------ fused link
@pytest.mark.parametrize('scheme,Nr,Nt,S,fnv', [('svd', 2, 2, 5, 0.0), ('svd', 3, 3, 4, 0.0), ('svd', 4, 4, 6, 0.0),
                                                ('gmd', 2, 2, 5, 0.0), ('gmd', 3, 3, 4, 0.03), ('gmd', 4, 4, 6, 0.03),
                                                ('mrt', 1, 1, 7, 0.0), ('mrt', 1, 2, 7, 0.0), ('mrt', 1, 4, 5, 0.0)])
def test_link_precoded_f64_vs_oracle(scheme, Nr, Nt, S, fnv):
    import torch
    from pyphysim_b200 import links
    pm, om = product_modem('qam', 16), oracle_modem('qam', 16)
    n, nv = 1500, 0.03
    layers = 1 if scheme == 'mrt' else Nt
    idx, H, nz = OL.draws_flat_mimo(SEED, np.arange(21, 21 + n), 4, Nr, Nt, S, S * layers)
    ref_hat, ref_dec = OL.precoded_flat(om, scheme, idx, H, nz, nv, fnv)
    kw = dict(scheme=scheme, Nr=Nr, Nt=Nt, num_symbols=S, filter_noise_var=fnv)
    cnt, hat, dec = links.link_precoded(pm, nv, n, dtype='f64', draws=(cuda(idx.astype(np.uint8)), cuda(H), cuda(nz)),
                                        want_idx=True, want_samples=True, **kw)
    # 1/S of a nearly singular H amplifies rounding: compare relative to the amplified scale
    assert_samples_close(_t(dec), ref_dec, 1e-7, scheme + ' f64')
    nbad = assert_decisions(_t(hat), ref_hat, om, ref_dec, exact=False, eps=1e-6, what=scheme + ' f64')
    ref_cnt = OL.counters(idx, ref_hat, 4)
    assert abs(int(cnt[0]) - int(ref_cnt[0])) <= nbad and cnt[2] == ref_cnt[2] == n * S * layers and cnt[3] == 4 * cnt[2]
    d = links.draw_flat_mimo(pm, n, Nr=Nr, Nt=Nt, num_symbols=S, n_data=S * layers, seed=SEED, first_unit=21, dtype='f64')
    assert np.array_equal(_t(d[0]), idx)
    c_s, hat_s = links.link_precoded(pm, nv, n, dtype='f64', draws=d, want_idx=True, **kw)
    c_f, hat_f = links.link_precoded(pm, nv, n, dtype='f64', seed=SEED, first_unit=21, want_idx=True, **kw)
    assert np.array_equal(c_s, c_f) and torch.equal(hat_s, hat_f)                         # fused == stream
    # float32 arithmetic (decomposition still in double)
    idx32, H32, nz32 = OL.draws_flat_mimo(SEED, np.arange(21, 21 + n), 4, Nr, Nt, S, S * layers, dtype=np.float32)
    ref_hat32, ref_dec32 = OL.precoded_flat(om, scheme, idx32, H32.astype(complex), nz32.astype(complex), nv, fnv)
    c32, hat32, dec32 = links.link_precoded(pm, nv, n, dtype='f32', draws=(cuda(idx32.astype(np.uint8)), cuda(H32), cuda(nz32)),
                                            want_idx=True, want_samples=True, **kw)
    assert_samples_close(_t(dec32), ref_dec32, 2e-4, scheme + ' f32')
    assert_decisions(_t(hat32), ref_hat32, om, ref_dec32, exact=False, eps=2e-3, what=scheme + ' f32')


def test_link_precoded_properties_and_errors():
    """Noiseless: every scheme decodes without error; GMD equalises the per-stream quality, so at high SNR
    its error rate is below SVD's (whose weakest eigenmode dominates); deterministic; argument checks."""
    from pyphysim_b200 import links
    pm = product_modem('qam', 16)
    n = 200000
    for scheme, Nr, Nt in (('svd', 4, 4), ('gmd', 4, 4), ('svd', 2, 2), ('gmd', 3, 3), ('mrt', 1, 4)):
        c = links.link_precoded(pm, 0.0, n, scheme=scheme, Nr=Nr, Nt=Nt, num_symbols=4)
        layers = 1 if scheme == 'mrt' else Nt
        assert c[0] == 0 and c[1] == 0 and c[2] == n * 4 * layers and c[3] == 4 * c[2]
    nv = 10 ** (-25 / 10)
    c_svd = links.link_precoded(pm, nv, n, scheme='svd', Nr=4, Nt=4, num_symbols=4)
    c_gmd = links.link_precoded(pm, nv, n, scheme='gmd', Nr=4, Nt=4, num_symbols=4)
    assert np.array_equal(c_svd, links.link_precoded(pm, nv, n, scheme='svd', Nr=4, Nt=4, num_symbols=4))
    assert 0 < c_gmd[0] < c_svd[0]
    c_mrt1 = links.link_precoded(pm, 0.1, n, scheme='mrt', Nr=1, Nt=1, num_symbols=4)
    c_mrt4 = links.link_precoded(pm, 0.1, n, scheme='mrt', Nr=1, Nt=4, num_symbols=4)
    assert c_mrt4[0] < c_mrt1[0]                                  # transmit diversity
    with pytest.raises(NotImplementedError):
        links.link_precoded(pm, 0.1, 10, scheme='svd', Nr=4, Nt=2)
    with pytest.raises(ValueError):
        links.link_precoded(pm, 0.1, 10, scheme='mrt', Nr=2, Nt=2)
    with pytest.raises(ValueError):
        links.link_precoded(pm, -0.1, 10, scheme='gmd', Nr=2, Nt=2)
    with pytest.raises(ValueError):
        links.link_precoded(pm, 0.1, 10, scheme='zf', Nr=2, Nt=2)
