"""Known-answer tables the reference's own tests hold for the hot path, restated
as literals and checked against the oracle (CPU, no reference needed).

Sources (relative to /root/reference): tests/modulators_package_test.py:45-70,
218-240, 402-420; tests/util_package_test.py:166-173, 337-380;
tests/channels_package_test.py:399-498; tests/mimo_package_test.py:610-637.
"""
import numpy as np
import pytest

from oracle import fading, mimo, modulators as md, ofdm, philox


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    def run(c, k):
        return [int(x) for x in philox.philox4x32_10(*[np.uint32(v) for v in c], *k)]
    assert run((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run((0xffffffff,) * 4, (0xffffffff,) * 2) == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_stream_layout():
    w = philox.words(123, 2, [5, 6], 10)
    w2 = philox.words(123, 2, [6], 8, first_word=4)
    assert np.array_equal(w[1, 4:10], w2[0, :6])
    assert not np.array_equal(philox.words(123, 1, [5], 4), philox.words(123, 2, [5], 4))
    idx = philox.data_indices(9, np.arange(1000), 16, 6)
    assert idx.min() == 0 and idx.max() == 63
    c = philox.cnormal(1, 2, np.arange(64), 4096)
    assert abs(np.mean(np.abs(c) ** 2) - 1) < 0.01 and abs(np.mean(c)) < 0.01
    u32 = philox.uniform(np.array([0, 1, 2 ** 32 - 1], dtype=np.uint32), np.float32)
    assert u32.dtype == np.float32 and u32[0] > 0 and u32[2] <= 1.0


def test_gray_tables():
    assert list(md.binary2gray(np.arange(8))) == [0, 1, 3, 2, 6, 7, 5, 4]
    assert np.array_equal(md.gray2binary(md.binary2gray(np.arange(256))), np.arange(256))


def test_level2bits_count_bits():
    assert [md.level2bits(n) for n in (1, 2, 3, 4, 5, 8, 9, 64, 256)] == [1, 1, 2, 2, 3, 3, 4, 6, 8]
    assert [md.int2bits(n) for n in (0, 1, 2, 3, 4, 255, 256)] == [1, 1, 2, 2, 3, 8, 9]
    with pytest.raises(ValueError):
        md.level2bits(0)
    n = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15])
    assert list(md.count_bits(n)) == [0, 1, 1, 2, 1, 2, 2, 3, 1, 2, 2, 3, 2, 3, 3, 4]
    a = np.random.RandomState(1).randint(0, 16, 20)
    b = np.random.RandomState(2).randint(0, 16, 20)
    assert md.count_bit_errors(a, b) == sum(bin(x ^ y).count('1') for x, y in zip(a, b))


def test_psk_tables():
    np.testing.assert_array_almost_equal(
        md.psk_constellation(4), np.array([1. + 0.j, 0. + 1.j, 0. - 1.j, -1. + 0.j]))
    s = 0.70710678
    np.testing.assert_array_almost_equal(
        md.psk_constellation(8),
        np.array([1, s + s * 1j, -s + s * 1j, 1j, s - s * 1j, -1j, -1, -s - s * 1j]))
    # setPhaseOffset rebuilds WITHOUT the Gray reorder (fundamental.py:459)
    np.testing.assert_array_almost_equal(
        md.psk_raw(4, np.pi / 4), np.array([s + s * 1j, -s + s * 1j, -s - s * 1j, s - s * 1j]))
    assert list(md.bpsk_constellation()) == [1, -1]


def test_qam_tables():
    a, b = 0.94868330, 0.31622777
    np.testing.assert_array_almost_equal(
        md.qam_constellation(4),
        np.array([-1 + 1j, 1 + 1j, -1 - 1j, 1 - 1j]) * 0.70710678)
    np.testing.assert_array_almost_equal(
        md.qam_constellation(16),
        np.array([-a + a * 1j, -b + a * 1j, a + a * 1j, b + a * 1j,
                  -a + b * 1j, -b + b * 1j, a + b * 1j, b + b * 1j,
                  -a - a * 1j, -b - a * 1j, a - a * 1j, b - a * 1j,
                  -a - b * 1j, -b - b * 1j, a - b * 1j, b - b * 1j]))
    q64 = md.qam_constellation(64)
    np.testing.assert_array_almost_equal(
        q64[:8].real, [-1.08012345, -0.77151675, -0.15430335, -0.46291005,
                       0.77151675, 1.08012345, 0.46291005, 0.15430335])
    assert abs(np.mean(np.abs(md.qam_constellation(256)) ** 2) - 1) < 1e-12
    for bad in (8, 32, 63):
        with pytest.raises(ValueError):
            md.qam_constellation(bad)


def test_modulate_errors_and_roundtrip():
    q = md.qam_constellation(16)
    with pytest.raises(ValueError):
        md.modulate(q, np.array([0, 16]))
    idx = np.arange(16)
    assert np.array_equal(md.demodulate(q, md.modulate(q, idx) + 0.01), idx)
    with pytest.raises(ValueError):
        md.bpsk_modulate(np.array([0, 2]))
    assert list(md.bpsk_demodulate(np.array([0.5 + 1j, -0.1 + 3j, 0 - 1j, 0 + 1j]))) == [0, 1, 1, 0]


def test_ofdm_bin_maps():
    # 52 of 64: data[0:26] -> bins 38..63, data[26:52] -> bins 1..26
    grid = np.zeros(64)
    grid[ofdm.used_subcarrier_indexes(64, 52)] = np.r_[1:53]
    expected = np.r_[0, 27:53, np.zeros(11), 1:27]
    assert np.array_equal(grid, expected)
    assert list(ofdm.used_subcarrier_indexes(16, 10)) == [11, 12, 13, 14, 15, 1, 2, 3, 4, 5]
    assert list(ofdm.used_subcarrier_indexes(16, 14)) == \
        [9, 10, 11, 12, 13, 14, 15, 1, 2, 3, 4, 5, 6, 7]
    # full allocation: data[0:half] -> bins fft/2.., data[half:] -> bins 0..
    assert list(ofdm.used_subcarrier_indexes(8, 8)) == [4, 5, 6, 7, 0, 1, 2, 3]
    assert ofdm.calc_zeropad(52, 60) == (8, 1)
    assert ofdm.power_scale(64, 16, 52) == 64.0 ** 2 / (52 + 16)
    for bad in ((64, 65, 52), (64, 16, 66), (64, 16, 51), (64, -1, 52)):
        with pytest.raises(ValueError):
            ofdm.check_parameters(*bad)


def test_ofdm_is_scaled_ifft_plus_cp():
    x = np.exp(1j * np.arange(52))
    t = ofdm.modulate(x, 64, 16, 52)
    grid = np.zeros(64, dtype=complex)
    grid[ofdm.used_subcarrier_indexes(64, 52)] = x
    body = np.sqrt(64.0 ** 2 / 68) * np.fft.ifft(grid)
    np.testing.assert_allclose(t, np.r_[body[-16:], body], atol=1e-14)
    np.testing.assert_allclose(ofdm.demodulate(t, 64, 16, 52), x, atol=1e-12)


def test_cost259_tu_discretisation():
    p, d = fading.discretize_profile(*fading.COST259_TU, 3.255e-08)
    assert list(d) == [0, 7, 16, 21, 27, 38, 40, 41, 47, 50, 56, 58, 60, 63, 66]
    assert abs(p.sum() - 1) < 1e-12
    _, d_ra = fading.discretize_profile(*fading.COST259_RA, 3.255e-08)
    _, d_ht = fading.discretize_profile(*fading.COST259_HT, 3.255e-08)
    assert (d_ra.size, d_ra[-1] + 1) == (10, 17)
    assert (d_ht.size, d_ht[-1] + 1) == (18, 554)


def test_jakes_clock():
    # tests/channels_package_test.py:244-259: after 1 + 100 samples the clock is 101 Ts
    Ts = 1e-3
    phi = np.zeros((4,))
    _, t1 = fading.jakes_samples(phi, phi, 5.0, Ts, 0.0, 1)
    _, t2 = fading.jakes_samples(phi, phi, 5.0, Ts, t1, 100)
    assert abs(t2 - 101 * Ts) < 1e-9


def test_alamouti_table_and_roundtrip():
    data = np.r_[0:16] + np.r_[0:16] * 1j
    exp0 = np.array([0 + 0j, -1 + 1j, 2 + 2j, -3 + 3j, 4 + 4j, -5 + 5j, 6 + 6j, -7 + 7j,
                     8 + 8j, -9 + 9j, 10 + 10j, -11 + 11j, 12 + 12j, -13 + 13j, 14 + 14j,
                     -15 + 15j])
    exp1 = np.array([1 + 1j, 0 - 0j, 3 + 3j, 2 - 2j, 5 + 5j, 4 - 4j, 7 + 7j, 6 - 6j, 9 + 9j,
                     8 - 8j, 11 + 11j, 10 - 10j, 13 + 13j, 12 - 12j, 15 + 15j, 14 - 14j])
    enc = mimo.alamouti_encode(data)
    np.testing.assert_array_almost_equal(enc, np.array([exp0, exp1]) / np.sqrt(2))
    H = philox.cnormal(3, 1, [0], 6)[0].reshape(3, 2)
    np.testing.assert_array_almost_equal(mimo.alamouti_decode(H @ enc, H), data)


def test_blast_roundtrip():
    H = philox.cnormal(3, 1, [1], 12)[0].reshape(4, 3)
    s = np.r_[0:15] + 0j
    x = mimo.blast_encode(s, 3)
    assert x.shape == (3, 5) and np.allclose(x[:, 0] * np.sqrt(3), [0, 1, 2])
    np.testing.assert_array_almost_equal(mimo.blast_decode(H @ x, H), s)
    np.testing.assert_array_almost_equal(mimo.blast_decode(H @ x, H, 1e-8), s, decimal=5)
    with pytest.raises(ValueError):
        mimo.blast_encode(np.zeros(7), 3)
