"""float32 (the arithmetic bench.py times) against float64 (the arithmetic held decision-exact to the oracle) on
the SAME device draws, at bench scale: how far the error counters drift and how close to a decision boundary
every disagreeing symbol is.  Device against device, so 1e5 frames take seconds."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (workload, frames, |count drift| / symbols, mismatch rate, worst margin of a mismatch)
# Bounds = ~3x the values measured on a B200 (profiles/parity_drift_r02.json): drift 10 / -7 / +6 symbols of
# 1.0e8 / 2.0e8 / 1.6e8, mismatch rates 6.3e-7 / 9.9e-7 / 4.5e-6, worst mismatch margins 1.7e-6 / 5.4e-6 / 1.6e-5
# against median decision margins of 0.146 / 0.169 / 0.059.
CASES = [
    ('c3_ofdm1024_qam64_siso_tdl', 100000, 5e-7, 2e-6, 1e-5),
    ('ofdm1024_qam64_mimo2x2_tdl', 100000, 5e-7, 3e-6, 2e-5),
    ('c5_ofdm2048_qam256_mimo4x4_tdl', 20000, 5e-7, 1.5e-5, 5e-5),
]


@pytest.mark.parametrize('wname,frames,drift_tol,rate_tol,margin_tol', CASES)
def test_f32_counters_against_f64_same_draws(wname, frames, drift_tol, rate_tol, margin_tol):
    import bench
    from pyphysim_b200.diagnostics import precision_drift
    w = bench.WORKLOADS[wname]
    link = bench.make_link(w)
    d = precision_drift(link, frames)
    print(wname, d)
    assert d['symbols'] == frames * link.n_data
    assert d['symbol_errors_f64'] > 1000                      # the counters are populated
    assert abs(d['symbol_count_drift']) <= max(2, drift_tol * d['symbols'])
    assert abs(d['bit_count_drift']) <= max(2, drift_tol * d['symbols'])
    assert d['mismatch_rate'] <= rate_tol
    assert d['worst_mismatch_margin'] <= margin_tol
