"""SURVEY.md §8f next-2, second half: the MuChannel / MuMimoChannel link grid.  CPU: the oracle restatement
against the fixture produced by the unmodified reference (tests/golden/make_golden_multiuser.py).  GPU: the
product classes (pyphysim_b200.channels.multiuser, all link arithmetic in libb200phy) against the same
fixture, with the fixture's Jakes phases written into every link's generator."""
import numpy as np
import pytest

from oracle import fading as OF
from oracle import multiuser as OMU

TOL = dict(rtol=1e-10, atol=1e-11)


def _grid(g, pre, shape, Fd, profile=None, powers=None, delays=None):
    Ts = float(g['a_Ts'])
    if profile is not None:
        powers, delays = OF.discretize_profile(profile[0], profile[1], Ts)
    phi = g[pre + 'phi'].reshape(shape + g[pre + 'phi'].shape[1:])[..., 0]
    psi = g[pre + 'psi'].reshape(shape + g[pre + 'psi'].shape[1:])[..., 0]
    return OMU.LinkGrid(phi, psi, Fd, Ts, powers, delays)


def test_oracle_grid_matches_reference(golden):
    g = golden('multiuser')
    grid = _grid(g, 'a_', (2, 3), 70.0, profile=OF.COST259_TU)
    grid.pathloss = g['a_pl']
    assert grid.delays.size == int(g['a_num_taps']) and grid.delays[-1] + 1 == int(g['a_pad'])
    y = grid.corrupt(g['a_x'])
    np.testing.assert_allclose(y[0], g['a_y0'], **TOL)
    np.testing.assert_allclose(y[1], g['a_y1'], **TOL)
    np.testing.assert_allclose(grid.last_taps[(1, 2)], g['a_ir12'], **TOL)          # includes sqrt(pathloss)
    y = grid.corrupt(g['a_x2'], freq=(64, None))
    np.testing.assert_allclose(y[0], g['a_y2_0'], **TOL)
    np.testing.assert_allclose(y[1], g['a_y2_1'], **TOL)
    y = grid.corrupt(g['a_x3'], freq=(64, g['a_car']))
    np.testing.assert_allclose(y[0], g['a_y3_0'], **TOL)
    np.testing.assert_allclose(y[1], g['a_y3_1'], **TOL)
    y = grid.corrupt(g['a_x4'], switched=True)
    assert len(y) == 3
    for k in range(3):
        np.testing.assert_allclose(y[k], g['a_y4_%d' % k], **TOL)
    # MIMO links, profile given as tap powers / delays
    Ts = float(g['a_Ts'])
    powers, delays = OF.discretize_profile(np.array([0.0, -6.0]), np.array([0.0, 3 * Ts]), Ts)
    grid = _grid(g, 'b_', (2, 2), 30.0, powers=powers, delays=delays)
    y = grid.corrupt([g['b_x0'], g['b_x1']])
    np.testing.assert_allclose(y[0], g['b_y0'], **TOL)
    np.testing.assert_allclose(y[1], g['b_y1'], **TOL)
    np.testing.assert_allclose(grid.last_taps[(0, 1)], g['b_ir01'], **TOL)


def _set_phases(mu, phi, psi):
    num_rx, num_tx = mu._su_siso_channels.shape
    for rx in range(num_rx):
        for tx in range(num_tx):
            gen = mu._su_siso_channels[rx, tx]._tdlchannel._fading_generator
            assert gen._phi_l.shape == phi[rx * num_tx + tx].shape
            gen._phi_l = phi[rx * num_tx + tx].copy()
            gen._psi_l = psi[rx * num_tx + tx].copy()


@pytest.mark.gpu
def test_mu_channel_classes_match_reference(golden):
    import torch
    from pyphysim_b200.channels import fading, fading_generators, multiuser
    g = golden('multiuser')
    Ts = float(g['a_Ts'])
    jakes = fading_generators.JakesSampleGenerator(Fd=70.0, Ts=Ts, L=8)
    mu = multiuser.MuChannel((2, 3), jakes, channel_profile=fading.COST259_TUx, Ts=Ts)
    _set_phases(mu, g['a_phi'], g['a_psi'])
    assert (mu.num_taps, mu.num_taps_with_padding, mu.switched_direction) == (int(g['a_num_taps']), int(g['a_pad']), False)
    assert 'MuChannel(shape=2x3, switched=False)' == repr(mu)
    mu.set_pathloss(g['a_pl'])
    np.testing.assert_array_equal(mu.pathloss_matrix, g['a_pl'])
    y = mu.corrupt_data(g['a_x'])
    assert y.dtype == object and y.shape == (2,)
    np.testing.assert_allclose(y[0], g['a_y0'], **TOL)
    np.testing.assert_allclose(y[1], g['a_y1'], **TOL)
    np.testing.assert_allclose(mu.get_last_impulse_response(1, 2).tap_values_sparse, g['a_ir12'], **TOL)
    y = mu.corrupt_data_in_freq_domain(g['a_x2'], 64)
    np.testing.assert_allclose(y[0], g['a_y2_0'], **TOL)
    np.testing.assert_allclose(y[1], g['a_y2_1'], **TOL)
    y = mu.corrupt_data_in_freq_domain(g['a_x3'], 64, g['a_car'])
    np.testing.assert_allclose(y[0], g['a_y3_0'], **TOL)
    np.testing.assert_allclose(y[1], g['a_y3_1'], **TOL)
    mu.switched_direction = True
    y = mu.corrupt_data(g['a_x4'])
    assert y.shape == (3,)
    for k in range(3):
        np.testing.assert_allclose(y[k], g['a_y4_%d' % k], **TOL)
    with pytest.raises(ValueError):
        mu.set_pathloss(np.full((2, 3), 1.5))
    # device tensors in -> device tensors out (no host round trip between the links)
    mu.switched_direction = False
    mu.set_pathloss(None)
    yt = mu.corrupt_data(torch.from_numpy(g['a_x']).cuda())
    assert yt[0].is_cuda and yt[0].shape == (g['a_x'].shape[1] + mu.num_taps_with_padding - 1,)

    jakes = fading_generators.JakesSampleGenerator(Fd=30.0, Ts=Ts, L=6)
    mm = multiuser.MuMimoChannel(2, 2, 3, jakes, tap_powers_dB=np.array([0.0, -6.0]),
                                 tap_delays=np.array([0.0, 3 * Ts]), Ts=Ts)
    _set_phases(mm, g['b_phi'], g['b_psi'])
    np.testing.assert_array_equal(mm.num_tx_antennas, g['b_ntx'])
    np.testing.assert_array_equal(mm.num_rx_antennas, g['b_nrx'])
    xs = np.empty(2, dtype=object)
    xs[0], xs[1] = g['b_x0'], g['b_x1']
    y = mm.corrupt_data(xs)
    np.testing.assert_allclose(y[0], g['b_y0'], **TOL)
    np.testing.assert_allclose(y[1], g['b_y1'], **TOL)
    np.testing.assert_allclose(mm.get_last_impulse_response(0, 1).tap_values_sparse, g['b_ir01'], **TOL)
    # default construction: independent flat Rayleigh links
    m0 = multiuser.MuChannel(3)
    y = m0.corrupt_data(np.ones((3, 10), dtype=complex))
    np.testing.assert_array_equal([v.size for v in y], g['c_len'])
    assert not np.allclose(y[0], y[1])
