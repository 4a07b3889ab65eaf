"""TdlChannel.corrupt_data_in_freq_domain: the oracle restatement against
the fixture made by the unmodified reference (CPU), and the GPU façade against the same fixture."""
import numpy as np
import pytest

from oracle import fading as ofading


def _oracle_taps(g, prefix, profile, Ts, Fd, num_blocks, fft, t0):
    p, d = ofading.discretize_profile(profile[0], profile[1], Ts)
    h, t1 = ofading.jakes_block_samples(g[prefix + '_phi'], g[prefix + '_psi'], Fd, Ts, t0, num_blocks, fft)
    return ofading.tdl_taps(h, p), d, t1


def test_oracle_freq_domain_matches_reference(golden):
    g = golden('freqdomain')
    Ts = float(g['s_Ts'])
    taps, d, t1 = _oracle_taps(g, 's', ofading.COST259_TU, Ts, 80.0, 5, 64, Ts)
    np.testing.assert_allclose(taps, g['s_taps'], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(ofading.tdl_corrupt_freq(g['s_x'], taps, d, 64, g['s_car']), g['s_y'],
                               rtol=1e-11, atol=1e-12)
    assert abs(t1 - float(g['s_t_end'])) < 1e-15
    taps2, d, _ = _oracle_taps(g, 's', ofading.COST259_TU, Ts, 80.0, 2, 64, t1)
    np.testing.assert_allclose(ofading.tdl_corrupt_freq(g['s_x'][:128], taps2, d, 64), g['s_y2'],
                               rtol=1e-11, atol=1e-12)
    Ts = float(g['m_Ts'])
    taps, d, _ = _oracle_taps(g, 'm', ofading.COST259_RA, Ts, 120.0, 4, 128, Ts)
    np.testing.assert_allclose(taps, g['m_taps'], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(ofading.tdl_corrupt_freq(g['m_x'], taps, d, 128, slice(10, 42)), g['m_y'],
                               rtol=1e-11, atol=1e-12)
    with pytest.raises(ValueError):
        ofading.tdl_corrupt_freq(g['m_x'][:, :100], taps, d, 128, slice(10, 42))


class QueueRS:
    def __init__(self):
        self.queue = []

    def rand(self, *shape):
        n = int(np.prod(shape))
        if self.queue and self.queue[0].size == n:
            return self.queue.pop(0).reshape(shape)
        return np.full(shape, 0.25)


@pytest.mark.gpu
def test_facade_freq_domain_matches_reference(golden):
    from pyphysim_b200.channels import fading
    from pyphysim_b200.channels.fading_generators import JakesSampleGenerator, RayleighSampleGenerator
    g = golden('freqdomain')
    tol = dict(rtol=1e-10, atol=1e-11)
    Ts = float(g['s_Ts'])
    rs = QueueRS()
    jakes = JakesSampleGenerator(Fd=80.0, Ts=Ts, L=20, RS=rs)
    prof = fading.COST259_TUx.get_discretize_profile(Ts)
    rs.queue = [g['s_phi'][..., None] / (2 * np.pi), g['s_psi'][..., None] / (2 * np.pi)]
    ch = fading.TdlChannel(jakes, prof)
    y = ch.corrupt_data_in_freq_domain(g['s_x'], 64, g['s_car'])
    np.testing.assert_allclose(y, g['s_y'], **tol)
    np.testing.assert_allclose(ch.get_last_impulse_response().tap_values_sparse, g['s_taps'], **tol)
    assert abs(jakes._current_time - float(g['s_t_end'])) < 1e-15
    y2 = ch.corrupt_data_in_freq_domain(g['s_x'][:128], 64)          # clock continues
    np.testing.assert_allclose(y2, g['s_y2'], **tol)
    np.testing.assert_allclose(ch.get_last_impulse_response().tap_values_sparse, g['s_taps2'], **tol)
    with pytest.raises(ValueError):
        ch.corrupt_data_in_freq_domain(g['s_x'][:100], 64)
    Ts = float(g['m_Ts'])
    rs = QueueRS()
    jakes = JakesSampleGenerator(Fd=120.0, Ts=Ts, L=16, shape=(3, 2), RS=rs)
    prof = fading.COST259_RAx.get_discretize_profile(Ts)
    rs.queue = [g['m_phi'][..., None] / (2 * np.pi), g['m_psi'][..., None] / (2 * np.pi)]
    ch = fading.TdlMimoChannel(jakes, prof)
    y = ch.corrupt_data_in_freq_domain(g['m_x'], 128, slice(10, 42))
    assert y.shape == (3, 128)
    np.testing.assert_allclose(y, g['m_y'], **tol)
    np.testing.assert_allclose(ch.get_last_impulse_response().tap_values_sparse, g['m_taps'], **tol)
    ch.switched_direction = True
    np.testing.assert_allclose(ch.corrupt_data_in_freq_domain(g['m_x3'], 128, slice(10, 42)), g['m_y3'], **tol)
    # Rayleigh generator: block-static multiplication by the per-block frequency response
    ch = fading.TdlChannel(RayleighSampleGenerator(), prof)
    x = np.ones(256, dtype=complex)
    y = ch.corrupt_data_in_freq_domain(x, 128)
    H = ch.get_last_impulse_response().get_freq_response(128)          # [128, 2 blocks]
    np.testing.assert_allclose(y, np.r_[H[:, 0], H[:, 1]], **tol)
