"""Size-independent properties of the oracle (hypothesis, CPU): the algebraic facts the GPU parity tests lean
on at full size - OFDM round trip for any legal parameter set, linearity of the time-varying TDL, the
circular-convolution identity behind the one-tap equaliser, demap(map(i)) == i under sub-half-distance noise,
Alamouti / Blast round trips, error counting symmetries."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import fading as F
from oracle import mimo as MI
from oracle import modulators as M
from oracle import ofdm as O

FAST = settings(max_examples=25, deadline=None)


def _cn(rng, *shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2.0)


@FAST
@given(lg=st.integers(3, 8), cp_frac=st.floats(0, 1), used_frac=st.floats(0.1, 1), n_sym=st.integers(1, 3),
       seed=st.integers(0, 2 ** 31))
def test_ofdm_roundtrip_any_parameters(lg, cp_frac, used_frac, n_sym, seed):
    fft = 1 << lg
    cp = int(cp_frac * fft)
    used = max(2, 2 * int(used_frac * fft / 2))
    O.check_parameters(fft, cp, used)
    rng = np.random.default_rng(seed)
    x = _cn(rng, n_sym * used)
    t = O.modulate(x, fft, cp, used)
    assert t.size == n_sym * (fft + cp)
    # the cyclic prefix is the tail of each symbol, and the frame has unit power scale convention
    sym = t.reshape(n_sym, fft + cp)
    np.testing.assert_allclose(sym[:, :cp], sym[:, fft:], atol=1e-12)
    np.testing.assert_allclose(O.demodulate(t, fft, cp, used), x, atol=1e-10)


@FAST
@given(n=st.integers(4, 64), ntaps=st.integers(1, 6), seed=st.integers(0, 2 ** 31), mimo=st.booleans())
def test_tdl_corrupt_is_linear_and_causal(n, ntaps, seed, mimo):
    rng = np.random.default_rng(seed)
    delays = np.sort(rng.choice(np.arange(0, 12), size=ntaps, replace=False))
    if mimo:
        taps = _cn(rng, ntaps, 3, 2, n)
        a, b = _cn(rng, 2, n), _cn(rng, 2, n)
    else:
        taps = _cn(rng, ntaps, n)
        a, b = _cn(rng, n), _cn(rng, n)
    ya, yb = F.tdl_corrupt(a, taps, delays), F.tdl_corrupt(b, taps, delays)
    np.testing.assert_allclose(F.tdl_corrupt(2.0 * a - 1j * b, taps, delays), 2.0 * ya - 1j * yb, atol=1e-10)
    assert ya.shape[-1] == n + int(delays[-1])
    # causal: zeroing the input from sample k on leaves the output before k + d_min untouched
    k = n // 2
    a2 = a.copy()
    a2[..., k:] = 0
    np.testing.assert_allclose(F.tdl_corrupt(a2, taps, delays)[..., :k + int(delays[0])],
                               ya[..., :k + int(delays[0])], atol=1e-12)


@FAST
@given(lg=st.integers(4, 7), ntaps=st.integers(1, 5), seed=st.integers(0, 2 ** 31))
def test_static_channel_with_cp_is_one_tap_per_subcarrier(lg, ntaps, seed):
    """With a channel that does not vary inside the symbol and a cyclic prefix at least as long as its memory,
    OFDM demodulation of the TDL output equals H_k times the sent symbols - the identity the one-tap equaliser
    (and its mean-taps restatement) rests on."""
    fft = 1 << lg
    rng = np.random.default_rng(seed)
    ntaps = min(ntaps, fft // 4)
    delays = np.sort(rng.choice(np.arange(0, fft // 4), size=ntaps, replace=False))
    cp = int(delays[-1]) + int(rng.integers(0, 3))
    x = _cn(rng, fft)
    t = O.modulate(x, fft, cp, fft)
    g = _cn(rng, ntaps)
    taps = np.repeat(g[:, None], t.size, axis=1)
    r = F.tdl_corrupt(t, taps, delays)[:t.size]
    y = O.demodulate(r, fft, cp, fft)
    H = O.mean_freq_response(taps, delays, fft, 1)
    np.testing.assert_allclose(H, O.mean_freq_response_reference(taps, delays, fft, 1), atol=1e-12)
    np.testing.assert_allclose(O.onetap_equalize(y, H, fft, fft), x, atol=1e-8)


@FAST
@given(kind=st.sampled_from(['qam4', 'qam16', 'qam64', 'qam256', 'psk8', 'psk16', 'qpsk']),
       seed=st.integers(0, 2 ** 31))
def test_demap_inverts_map_under_sub_half_distance_noise(kind, seed):
    tab = {'qam4': lambda: M.qam_constellation(4), 'qam16': lambda: M.qam_constellation(16),
           'qam64': lambda: M.qam_constellation(64), 'qam256': lambda: M.qam_constellation(256),
           'psk8': lambda: M.psk_constellation(8), 'psk16': lambda: M.psk_constellation(16),
           'qpsk': M.qpsk_constellation}[kind]()
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, tab.size, 500)
    d = np.abs(tab[:, None] - tab[None, :])
    dmin = d[d > 0].min()
    noise = 0.49 * dmin * rng.uniform(0, 1, idx.size) * np.exp(2j * np.pi * rng.uniform(0, 1, idx.size))
    assert np.array_equal(M.demodulate(tab, M.modulate(tab, idx) + noise), idx)


@FAST
@given(a=st.lists(st.integers(0, 255), min_size=1, max_size=50), seed=st.integers(0, 2 ** 31))
def test_bit_error_count_is_a_metric(a, seed):
    a = np.array(a)
    rng = np.random.default_rng(seed)
    b, c = rng.integers(0, 256, a.size), rng.integers(0, 256, a.size)
    dab, dbc, dac = M.count_bit_errors(a, b), M.count_bit_errors(b, c), M.count_bit_errors(a, c)
    assert M.count_bit_errors(a, a) == 0 and dab == M.count_bit_errors(b, a) and dac <= dab + dbc
    assert dab == sum(bin(int(x) ^ int(y)).count('1') for x, y in zip(a, b))


@FAST
@given(nr=st.integers(1, 4), ncw=st.integers(1, 6), seed=st.integers(0, 2 ** 31))
def test_alamouti_roundtrip_noise_free(nr, ncw, seed):
    rng = np.random.default_rng(seed)
    s = _cn(rng, 2 * ncw)
    H = _cn(rng, nr, 2)
    np.testing.assert_allclose(MI.alamouti_decode(H @ MI.alamouti_encode(s), H), s, atol=1e-10)


@FAST
@given(nt=st.integers(1, 4), extra=st.integers(0, 2), t=st.integers(1, 5), seed=st.integers(0, 2 ** 31))
def test_blast_zf_roundtrip_and_mmse_shrinks(nt, extra, t, seed):
    rng = np.random.default_rng(seed)
    nr = nt + extra
    H = _cn(rng, nr, nt)
    x = _cn(rng, nt * t)
    y = H @ MI.blast_encode(x, nt)
    np.testing.assert_allclose(MI.blast_decode(y, H), x, atol=1e-6 * np.linalg.cond(H))
    z = MI.blast_decode(y, H, noise_var=0.5)
    assert np.linalg.norm(z) <= np.linalg.norm(x) * (1 + 1e-9)     # MMSE regularisation never amplifies
