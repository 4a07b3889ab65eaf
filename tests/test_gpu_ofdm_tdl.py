"""GPU parity of the fused OFDM-over-Jakes/TDL link (configs C3, C5 and the 2x2 headline workload)
against the oracle and the golden fixtures made from the unmodified reference."""
import numpy as np
import pytest

from oracle import fading
from oracle import links as OL
from oracle import modulators as md

from gpu_util import (assert_decisions, assert_samples_close, cuda, eq_tolerance, oracle_modem, product_modem)

pytestmark = pytest.mark.gpu
SEED = 0xC0FFEE


def _t(x):
    return x.cpu().numpy()


def make_pair(kind, M, fft, cp, used, n_sym=1, Nr=1, Nt=1, profile=fading.COST259_TU, Fd=10.0, L=20,
              snr_dB=20.0, fnv=None, dtype='f64', jakes='auto', Ts=None, t0=None):
    """(oracle config, product link) describing the same link."""
    from pyphysim_b200 import links
    om = oracle_modem(kind, M)
    nv = 1 / md.dB2Linear(snr_dB)
    cfg = OL.OfdmTdlConfig(om, fft, cp, used, n_sym=n_sym, Nr=Nr, Nt=Nt, profile=profile, Fd=Fd, L=L,
                           noise_var=nv, filter_noise_var=fnv, Ts=Ts, t0=t0)
    link = links.OfdmTdlLink(product_modem(kind, M), fft, cp, used, num_ofdm_symbols=n_sym, Nr=Nr, Nt=Nt,
                             tap_powers_linear=cfg.tap_powers, tap_delays=cfg.delays, Fd=Fd, Ts=cfg.Ts,
                             L=L, t0=cfg.t0, noise_var=nv, filter_noise_var=fnv, dtype=dtype,
                             jakes_mode=jakes, seed=SEED)
    return cfg, link


REL_F32 = 1e-5      # BASELINE.json north_star: complex sample values within 1e-5 relative
DEC_REL = 2e-6      # decisions: a mismatch must sit within 4x the bound at this (measured-error-sized) level


def run_stream_vs_oracle(cfg, link, units, exact, rel, eps=2e-4):
    """Host Philox draws -> (oracle, device stream mode).  Three comparisons:
      * the demodulated rx samples BEFORE detection (OFDM.demodulate output per rx antenna) at `rel`
        (float32: 1e-5 for every antenna shape) of max(|ref|, rms);
      * the equalised symbols: float64 at `rel` of max(|ref|, rms); float32 at the per-subcarrier first-order
        bound rel * ||G_k|| (||y_k|| + ||H_k|| ||z_k||) (gpu_util.eq_tolerance; SISO: G_k = 1 / H_k) — the
        receive filter / one-tap division amplifies the 1e-5 of its inputs by its norm, which is what makes
        a deep fade (|H_k| ~ 1e-3) look like a 5e-5 "relative" error on a symbol of magnitude 50;
      * the decisions: identical (exact), or any mismatch has an oracle margin below `eps`
        (float32: below 4x that symbol's sample bound at DEC_REL, i.e. a few times the measured error)."""
    f32 = link.dtype == 0
    npdt = np.float32 if f32 else np.float64
    idx, phi, psi, noise = OL.draws_ofdm_tdl(cfg, SEED, units, dtype=npdt)
    n = len(units)
    draws = (cuda(idx.astype(np.uint8)), cuda(phi.reshape(n, -1)), cuda(psi.reshape(n, -1)), cuda(noise))
    cnt, hat, eq, rx = link.run(n, first_unit=int(units[0]), draws=draws, want_idx=True, want_eq=True,
                                want_rx=True)
    ref_hat = np.empty_like(idx)
    ref_eq = np.empty(idx.shape, dtype=complex)
    ref_rx = np.empty((n, cfg.Nr, cfg.n_sym * cfg.used), dtype=complex)
    tol_eq = np.empty(idx.shape) if f32 else None
    for u in range(n):
        ref_hat[u], det = OL.ofdm_tdl_frame(cfg, idx[u], phi[u].astype(np.float64), psi[u].astype(np.float64),
                                            noise[u].astype(np.complex128), detail=True)
        ref_eq[u] = det['eq']
        ref_rx[u] = det['Y']
        if tol_eq is not None:
            tol_eq[u] = eq_tolerance(det, rel, cfg.Nr, cfg.Nt)
    worst_rx = assert_samples_close(_t(rx), ref_rx, rel, 'rx samples before detection')
    worst_eq = assert_samples_close(_t(eq), ref_eq, rel, 'equalised symbols', tol=tol_eq)
    if tol_eq is not None:
        eps = 4.0 * tol_eq * (DEC_REL / rel)
    nbad = assert_decisions(_t(hat), ref_hat, cfg.modem, ref_eq, exact=exact, eps=eps)
    ref_cnt = OL.counters(idx, ref_hat, cfg.modem.bits)
    assert abs(int(cnt[0]) - int(ref_cnt[0])) <= nbad and abs(int(cnt[1]) - int(ref_cnt[1])) <= 8 * nbad
    assert cnt[2] == ref_cnt[2] and cnt[3] == ref_cnt[3]
    print('parity %dx%d fft %d %s: %d frames, worst rx err/tol %.3f, worst eq err/tol %.3f, %d boundary mismatches'
          % (cfg.Nr, cfg.Nt, cfg.fft, 'f32' if f32 else 'f64', n, worst_rx, worst_eq, nbad))
    return cnt, ref_cnt


# ------------------------------------------------------------------ draws
def test_ofdm_draws_match_host_philox():
    cfg, link = make_pair('qam', 64, 128, 16, 100, n_sym=2, Nr=2, Nt=2, dtype='f64')
    idx, phi, psi, noise = link.draw(40, 3)
    hi, hphi, hpsi, hn = OL.draws_ofdm_tdl(cfg, SEED, np.arange(40, 43))
    assert np.array_equal(_t(idx), hi)
    assert np.array_equal(_t(phi), hphi.reshape(3, -1))            # phases: bit exact
    assert np.array_equal(_t(psi), hpsi.reshape(3, -1))
    np.testing.assert_allclose(_t(noise), hn, atol=1e-13)
    cfg, link = make_pair('qam', 64, 128, 16, 100, dtype='f32')
    idx, phi, psi, noise = link.draw(0, 2)
    hi, hphi, hpsi, hn = OL.draws_ofdm_tdl(cfg, SEED, np.arange(2), dtype=np.float32)
    assert np.array_equal(_t(phi), hphi.reshape(2, -1)) and np.array_equal(_t(psi), hpsi.reshape(2, -1))
    np.testing.assert_allclose(_t(noise), hn, atol=2e-6)


# ------------------------------------------------------------------ golden fixtures (reference output)
def test_c3_golden_full_size(golden):
    g = golden('links')
    for jakes in ('recurrence', 'auto'):
        cfg, link = make_pair('qam', 64, 1024, 72, 1024, dtype='f64', jakes=jakes)
        idx, phi, psi, noise = OL.draws_ofdm_tdl(cfg, SEED, g['c3_units'])
        draws = (cuda(idx.astype(np.uint8)), cuda(phi.reshape(2, -1)), cuda(psi.reshape(2, -1)), cuda(noise))
        cnt, hat, eq = link.run(2, draws=draws, want_idx=True, want_eq=True)
        assert_samples_close(_t(eq), g['c3_eq'], 1e-9, 'C3 vs reference (%s)' % jakes)
        assert np.array_equal(_t(hat), g['c3_hat'])
        assert np.array_equal(cnt, OL.counters(idx, g['c3_hat'], 6)) and cnt[0] > 0


def test_mimo2x2_golden(golden):
    g = golden('links')
    cfg, link = make_pair('qam', 16, 256, 18, 200, n_sym=2, Nr=2, Nt=2, Fd=300.0, snr_dB=25.0,
                          dtype='f64', jakes='recurrence')
    idx, phi, psi, noise = OL.draws_ofdm_tdl(cfg, SEED, g['m2_units'])
    draws = (cuda(idx.astype(np.uint8)), cuda(phi.reshape(2, -1)), cuda(psi.reshape(2, -1)), cuda(noise))
    cnt, hat, eq = link.run(2, first_unit=10, draws=draws, want_idx=True, want_eq=True)
    assert_samples_close(_t(eq), g['m2_eq'], 1e-8, '2x2 vs reference')
    assert np.array_equal(_t(hat), g['m2_hat'])


# ------------------------------------------------------------------ oracle parity, f64
@pytest.mark.parametrize('jakes', ['recurrence', 'poly'])
def test_c3_f64(jakes):
    Fd = 10.0 if jakes == 'recurrence' else 2.0
    cfg, link = make_pair('qam', 64, 1024, 72, 1024, dtype='f64', jakes=jakes, Fd=Fd)
    run_stream_vs_oracle(cfg, link, np.arange(100, 104), exact=True, rel=1e-9)


@pytest.mark.parametrize('case', [
    dict(kind='qam', M=16, fft=64, cp=16, used=52, n_sym=3, Fd=500.0, profile=fading.COST259_RA, Ts=2e-7),
    dict(kind='psk', M=8, fft=128, cp=0, used=60, n_sym=2, Fd=50.0, Ts=1e-7),
    dict(kind='bpsk', M=2, fft=256, cp=32, used=256, n_sym=1, Fd=1000.0, L=8, Ts=1e-7),
    dict(kind='qam', M=256, fft=2048, cp=144, used=1200, n_sym=1, Fd=10.0),
    dict(kind='qam', M=64, fft=512, cp=52, used=300, n_sym=2, Fd=5.0, Ts=2.0e-7, L=33),
    dict(kind='qam', M=16, fft=256, cp=200, used=100, n_sym=2, Fd=100.0, profile=fading.COST259_HT, Ts=1e-7),
])
def test_siso_shapes_f64(case):
    cfg, link = make_pair(dtype='f64', **case)
    run_stream_vs_oracle(cfg, link, np.arange(2), exact=True, rel=1e-9)


@pytest.mark.parametrize('case', [
    dict(kind='qam', M=64, fft=1024, cp=72, used=1024, Nr=2, Nt=2, snr_dB=25.0),           # headline workload
    dict(kind='qam', M=16, fft=128, cp=16, used=100, n_sym=2, Nr=2, Nt=2, Fd=400.0, fnv=0.0, Ts=1e-7),  # ZF
    dict(kind='qam', M=16, fft=128, cp=16, used=100, n_sym=2, Nr=2, Nt=1, Fd=400.0, Ts=1e-7),
    dict(kind='qam', M=16, fft=256, cp=18, used=200, Nr=4, Nt=2, snr_dB=15.0),
    dict(kind='qam', M=256, fft=2048, cp=144, used=2048, Nr=4, Nt=4, snr_dB=30.0),          # C5
    dict(kind='qam', M=16, fft=128, cp=16, used=100, n_sym=2, Nr=3, Nt=1, Fd=300.0, Ts=1e-7),
    dict(kind='qam', M=16, fft=128, cp=16, used=100, Nr=3, Nt=2, snr_dB=18.0, Fd=300.0, Ts=1e-7),
    dict(kind='qam', M=64, fft=256, cp=18, used=200, Nr=3, Nt=3, snr_dB=28.0),
    dict(kind='psk', M=8, fft=128, cp=16, used=100, Nr=4, Nt=1, Fd=300.0, Ts=1e-7),
    dict(kind='qam', M=16, fft=256, cp=18, used=200, n_sym=2, Nr=4, Nt=3, snr_dB=24.0),
])
def test_mimo_shapes_f64(case):
    cfg, link = make_pair(dtype='f64', **case)
    # the normal-equation solve in double vs LAPACK pinv/solve: 1e-7 at ill-conditioned subcarriers
    run_stream_vs_oracle(cfg, link, np.arange(7, 9), exact=False, rel=1e-7, eps=1e-6)


# ------------------------------------------------------------------ f32 (throughput arithmetic)
@pytest.mark.parametrize('case,jakes', [
    (dict(kind='qam', M=64, fft=1024, cp=72, used=1024), 'auto'),
    (dict(kind='qam', M=64, fft=1024, cp=72, used=1024), 'recurrence'),
    (dict(kind='qam', M=64, fft=1024, cp=72, used=1024, Nr=2, Nt=2, snr_dB=25.0), 'auto'),
    (dict(kind='qam', M=256, fft=2048, cp=144, used=2048, Nr=4, Nt=4, snr_dB=30.0), 'auto'),
    (dict(kind='qam', M=16, fft=128, cp=16, used=100, n_sym=3, Fd=800.0, Ts=1e-7), 'auto'),
    (dict(kind='qam', M=16, fft=256, cp=18, used=200, Nr=3, Nt=3, snr_dB=24.0), 'auto'),
    (dict(kind='qam', M=16, fft=256, cp=18, used=200, Nr=4, Nt=3, snr_dB=24.0), 'auto'),
])
def test_f32_within_tolerance(case, jakes):
    cfg, link = make_pair(dtype='f32', jakes=jakes, **case)
    # 1e-5 relative on samples (BASELINE north_star) at every antenna shape: rx samples before detection and SISO
    # equalised symbols directly, MIMO equalised symbols through the per-subcarrier ||G_k|| bound.  SISO decision
    # mismatches must sit within 4x the worst equalised-sample error (1e-5 x |z| <~ 1.5) of a boundary.
    run_stream_vs_oracle(cfg, link, np.arange(50, 53), exact=False, rel=REL_F32)


@pytest.mark.parametrize('case,nframes', [
    (dict(kind='qam', M=64, fft=1024, cp=72, used=1024), 64),                                       # C3
    (dict(kind='qam', M=64, fft=1024, cp=72, used=1024, Nr=2, Nt=2, snr_dB=25.0), 64),             # headline
    (dict(kind='qam', M=256, fft=2048, cp=144, used=2048, Nr=4, Nt=4, snr_dB=30.0), 64),           # C5
])
def test_f32_bench_kernels_64_frames(case, nframes):
    """The three float32 kernels bench.py times (frame-pair, antenna-pair 2x2, antenna-pair 4x4) against the
    float64 oracle on 64 full-size frames each (131 k / 131 k / 524 k symbols)."""
    cfg, link = make_pair(dtype='f32', **case)
    run_stream_vs_oracle(cfg, link, np.arange(1000, 1000 + nframes), exact=False, rel=REL_F32)


@pytest.mark.parametrize('case', [
    dict(kind='qam', M=64, fft=1024, cp=72, used=600),                          # LTE 10 MHz numerology, guard band
    dict(kind='qam', M=16, fft=2048, cp=144, used=1200),                        # LTE 20 MHz
    dict(kind='qam', M=64, fft=1024, cp=72, used=600, Nr=2, Nt=2, snr_dB=25.0),
    dict(kind='qam', M=16, fft=2048, cp=144, used=1200, Nr=2, Nt=2, snr_dB=22.0),
])
def test_f32_pair_kernels_guard_band(case):
    """used < fft in float32 at fft 1024 / 2048: the run-time-shape (LGF = 0) instantiations of the frame-pair and
    antenna-pair kernels with unused bins (pos_of = -1), against the oracle and against the generic kernel."""
    import torch
    from pyphysim_b200 import links
    cfg, pair = make_pair(dtype='f32', **case)
    run_stream_vs_oracle(cfg, pair, np.arange(20, 23), exact=False, rel=REL_F32)
    gen = links.OfdmTdlLink(pair.modulator, cfg.fft, cfg.cp, cfg.used, Nr=cfg.Nr, Nt=cfg.Nt,
                            tap_powers_linear=cfg.tap_powers, tap_delays=cfg.delays, Fd=10.0, Ts=cfg.Ts, L=20,
                            t0=cfg.t0, noise_var=cfg.noise_var, dtype='f32', seed=SEED, use_pair_kernel=False)
    draws = pair.draw(20, 5)
    c_p, hat_p, eq_p = pair.run(5, first_unit=20, draws=draws, want_idx=True, want_eq=True)
    c_g, hat_g, eq_g = gen.run(5, first_unit=20, draws=draws, want_idx=True, want_eq=True)
    assert_samples_close(_t(eq_p), _t(eq_g), 1e-4 if cfg.mimo else 3e-5, 'pair vs generic (guard band)')
    assert c_p[2] == c_g[2] == 5 * cfg.n_data
    c_f, hat_f = pair.run(5, first_unit=20, want_idx=True)
    assert np.array_equal(c_f, c_p) and torch.equal(hat_f, hat_p)


# ------------------------------------------------------------------ fused mode, sharding, properties
@pytest.mark.parametrize('dtype', ['f32', 'f64'])
@pytest.mark.parametrize('ant', [(1, 1), (2, 2)])
def test_fused_equals_stream_on_device_draws(dtype, ant):
    import torch
    cfg, link = make_pair('qam', 64, 256, 18, 200, n_sym=2, Nr=ant[0], Nt=ant[1], dtype=dtype, Fd=200.0)
    n = 37
    draws = link.draw(1000, n)
    c_s, hat_s = link.run(n, first_unit=1000, draws=draws, want_idx=True)
    c_f, hat_f = link.run(n, first_unit=1000, want_idx=True)
    assert np.array_equal(c_s, c_f) and torch.equal(hat_s, hat_f)
    acc = torch.zeros(4, dtype=torch.int64, device='cuda')
    for a, b in ((0, 5), (5, 6), (6, n)):
        link.run(b - a, first_unit=1000 + a, counters=acc)
    assert np.array_equal(_t(acc), c_f)                      # invariant to batch split / GPU sharding
    c_h, hat_h = link.run_host(n, first_unit=1000, draws=tuple(t.cpu().pin_memory() for t in draws),
                               want_idx=True)
    assert np.array_equal(c_h, c_f) and np.array_equal(hat_h.numpy(), _t(hat_f))


@pytest.mark.parametrize('ant,fft,cp,nsym', [((2, 2), 1024, 72, 2), ((4, 4), 2048, 144, 1), ((4, 2), 1024, 72, 1)])
def test_pair_kernel_vs_generic_kernel(ant, fft, cp, nsym):
    """The antenna-pair FFMA2 kernel against the generic kernel (same draws, float): same decisions
    up to boundary symbols, samples within 1e-5; and its own fused mode == its stream mode."""
    import torch
    from pyphysim_b200 import links
    M = 64 if ant == (2, 2) else 16
    cfg, pair = make_pair('qam', M, fft, cp, fft, n_sym=nsym, Nr=ant[0], Nt=ant[1], dtype='f32', snr_dB=22.0)
    gen = links.OfdmTdlLink(pair.modulator, fft, cp, fft, num_ofdm_symbols=nsym, Nr=ant[0], Nt=ant[1],
                            tap_powers_linear=cfg.tap_powers, tap_delays=cfg.delays, Fd=10.0, Ts=cfg.Ts, L=20,
                            t0=cfg.t0, noise_var=cfg.noise_var, dtype='f32', seed=SEED, use_pair_kernel=False)
    n = 12
    draws = pair.draw(300, n)
    c_p, hat_p, eq_p = pair.run(n, first_unit=300, draws=draws, want_idx=True, want_eq=True)
    c_g, hat_g, eq_g = gen.run(n, first_unit=300, draws=draws, want_idx=True, want_eq=True)
    assert_samples_close(_t(eq_p), _t(eq_g), 2e-5, 'pair vs generic')
    nbad = assert_decisions(_t(hat_p), _t(hat_g), cfg.modem, _t(eq_g).astype(complex), exact=False, eps=2e-3)
    assert abs(int(c_p[0]) - int(c_g[0])) <= nbad and c_p[2] == c_g[2]
    c_f, hat_f = pair.run(n, first_unit=300, want_idx=True)
    assert np.array_equal(c_f, c_p) and torch.equal(hat_f, hat_p)            # fused == stream, pair kernel
    # oracle, one frame (full-size frames are slow in NumPy)
    run_stream_vs_oracle(cfg, pair, np.arange(300, 301), exact=False, rel=REL_F32)


@pytest.mark.parametrize('mod,M,fft,cp,nsym,n', [('qam', 64, 1024, 72, 1, 13), ('psk', 8, 2048, 144, 2, 6),
                                                 ('qam', 16, 1024, 0, 3, 2)])
def test_frame_pair_kernel_vs_generic_kernel(mod, M, fft, cp, nsym, n):
    """SISO: the two-frames-per-CTA FFMA2 kernel (odd batch: the last frame runs in a pair with a masked lane)
    against the generic kernel on the same draws, its fused mode against its stream mode, and the
    oracle on the first two frames."""
    import torch
    from pyphysim_b200 import links
    cfg, pair = make_pair(mod, M, fft, cp, fft, n_sym=nsym, dtype='f32', snr_dB=24.0)
    gen = links.OfdmTdlLink(pair.modulator, fft, cp, fft, num_ofdm_symbols=nsym, Nr=1, Nt=1,
                            tap_powers_linear=cfg.tap_powers, tap_delays=cfg.delays, Fd=10.0, Ts=cfg.Ts, L=20,
                            t0=cfg.t0, noise_var=cfg.noise_var, dtype='f32', seed=SEED, use_pair_kernel=False)
    draws = pair.draw(500, n)
    c_p, hat_p, eq_p = pair.run(n, first_unit=500, draws=draws, want_idx=True, want_eq=True)
    c_g, hat_g, eq_g = gen.run(n, first_unit=500, draws=draws, want_idx=True, want_eq=True)
    # two float32 kernels with different summation orders (class-sorted H_k, FFT passes fused through registers):
    # 3e-5 of max(|ref|, rms) between them - the cp = 0 case (no cyclic prefix: ISI, equalised rms 4.6) sits at
    # 2.1e-5 in its deepest fade; each kernel is held to the float64 oracle separately below
    assert_samples_close(_t(eq_p), _t(eq_g), 3e-5, 'frame-pair vs generic')
    nbad = assert_decisions(_t(hat_p), _t(hat_g), cfg.modem, _t(eq_g).astype(complex), exact=False, eps=2e-3)
    assert abs(int(c_p[0]) - int(c_g[0])) <= nbad and c_p[2] == c_g[2] == n * nsym * fft
    c_f, hat_f = pair.run(n, first_unit=500, want_idx=True)
    assert np.array_equal(c_f, c_p) and torch.equal(hat_f, hat_p)
    c_n = pair.run(n, first_unit=500)                                        # counters only, no outputs
    assert np.array_equal(c_n, c_p)
    run_stream_vs_oracle(cfg, pair, np.arange(500, 502), exact=False, rel=REL_F32)


@pytest.mark.parametrize('ant,fft,cp,M,n', [((1, 1), 1024, 72, 64, 3001), ((2, 2), 1024, 72, 64, 3000),
                                            ((4, 4), 2048, 144, 256, 400), ((4, 2), 1024, 72, 16, 600)])
def test_fused_equals_stream_at_scale(ant, fft, cp, M, n):
    """Every float32 fast kernel (frame-pair, antenna-pair 2x2 / 4x2 / 4x4): the Monte Carlo mode and the stream mode
    fed with the device's own draws give identical decisions AND identical equalised samples on thousands of
    frames (millions of symbols).  Both modes apply sigma to the unit-variance normals with the same single FFMA2."""
    import torch
    cfg, link = make_pair('qam', M, fft, cp, fft, Nr=ant[0], Nt=ant[1], dtype='f32', snr_dB=24.0)
    first = 123
    draws = link.draw(first, n)
    c_s, hat_s, eq_s = link.run(n, first_unit=first, draws=draws, want_idx=True, want_eq=True)
    c_f, hat_f, eq_f = link.run(n, first_unit=first, want_idx=True, want_eq=True)
    assert np.array_equal(c_s, c_f) and torch.equal(hat_s, hat_f)
    assert torch.equal(eq_s.view(torch.float32), eq_f.view(torch.float32))


def test_tensor_core_channel_matrices_variant():
    """The opt-in tcgen05 variant of the headline kernel (per-subcarrier channel matrices H_k as one 3xTF32 tensor-core
    tile per frame, ofdm_tdl_pair.cuh template parameter TC): same oracle tolerances as the default kernel, fused ==
    stream bit for bit, and decisions within a few boundary symbols of the CUDA-core kernel on the same frames."""
    import torch
    from pyphysim_b200 import _lib
    cfg, link = make_pair('qam', 64, 1024, 72, 1024, Nr=2, Nt=2, dtype='f32', snr_dB=25.0)
    link.params.reserved |= 2
    run_stream_vs_oracle(cfg, link, np.arange(1000, 1016), exact=False, rel=REL_F32)
    assert _lib.load().b200phy_last_kernel().decode().endswith(',10,1>')          # the tensor-core instantiation ran
    n, first = 2000, 77
    draws = link.draw(first, n)
    c_s, hat_s = link.run(n, first_unit=first, draws=draws, want_idx=True)
    c_f, hat_f = link.run(n, first_unit=first, want_idx=True)
    assert np.array_equal(c_s, c_f) and torch.equal(hat_s, hat_f)                  # fused == stream bit for bit
    link.params.reserved &= ~2
    c_c, hat_c = link.run(n, first_unit=first, draws=draws, want_idx=True)
    assert not _lib.load().b200phy_last_kernel().decode().endswith(',10,1>')
    differing = int((hat_c != hat_s).sum())
    assert differing <= max(4, 5e-6 * hat_s.numel()), differing                   # 3xTF32 vs FFMA2: boundary symbols only
    assert abs(int(c_c[0]) - int(c_s[0])) <= differing


@pytest.mark.parametrize('ant', [(1, 1), (2, 2)])
@pytest.mark.parametrize('shift,cp', [(0, 72), (1, 72), (1, 0)])
def test_tma_input_pipeline_any_row_alignment(ant, shift, cp):
    """Stream mode of the two float32 kernels whose noise rows and phases arrive by TMA bulk copies
    (cp.async.bulk + mbarrier, ofdm_tdl_pair.cuh / ofdm_tdl_fpair.cuh): the rows of consecutive frames start
    alternately 0 and 8 bytes off a 16-byte boundary (odd row length), and a sliced noise / phase tensor moves
    everything by one more element — every combination must give the same result bit for bit."""
    import torch
    # (cp = 0 with a shifted tensor: the first row would have to be copied from before the tensor -> falls back to cp.async)
    cfg, link = make_pair('qam', 64, 1024, cp, 1024, Nr=ant[0], Nt=ant[1], dtype='f32', snr_dB=24.0)
    n, first = 9, 40
    idx, phi, psi, noise = link.draw(first, n)
    c_f, hat_f = link.run(n, first_unit=first, want_idx=True)

    def shifted(t):
        flat = torch.empty(t.numel() + 1, dtype=t.dtype, device=t.device)
        flat[shift:shift + t.numel()] = t.reshape(-1)
        return flat[shift:shift + t.numel()].view(t.shape)

    # reference: the same kernel on the tensors as drawn (16-byte aligned bases)
    c_a, hat_a = link.run(n, first_unit=first, draws=(idx, phi, psi, noise), want_idx=True)
    c_s, hat_s = link.run(n, first_unit=first, draws=(idx, shifted(phi), shifted(psi), shifted(noise)), want_idx=True)
    assert np.array_equal(c_s, c_a) and torch.equal(hat_s, hat_a)        # alignment must not change a single bit
    # every frame on its own: each row parity in the first landing slot
    acc = torch.zeros(4, dtype=torch.int64, device='cuda')
    for u in range(n):
        link.run(1, first_unit=first + u, draws=(idx[u:u + 1], phi[u:u + 1], psi[u:u + 1], shifted(noise)[u:u + 1]),
                 counters=acc)
    assert np.array_equal(_t(acc), c_a)
    assert np.array_equal(c_f, c_a) and torch.equal(hat_f, hat_a)        # and equals the fused-RNG mode bit for bit


@pytest.mark.parametrize('n,first', [(1, 0), (7, 0), (7, 3), (20, 11)])
def test_frame_pair_kernel_is_invariant_to_batch_and_shard_boundaries(n, first):
    """a frame's decisions must not depend on which lane / pair / batch it ran in.  Odd batches, odd
    shard starts, a batch of one: the per-frame outputs equal those of one big even batch, bit for bit."""
    import torch
    cfg, link = make_pair('qam', 64, 1024, 72, 1024, dtype='f32', snr_dB=18.0)
    big_c, big_hat, big_eq = link.run(32, first_unit=0, want_idx=True, want_eq=True)
    c, hat, eq = link.run(n, first_unit=first, want_idx=True, want_eq=True)
    assert torch.equal(hat, big_hat[first:first + n]) and torch.equal(eq, big_eq[first:first + n])
    assert c[2] == n * 1024
    draws = link.draw(first, n)                                    # stream mode, same frames
    c_s, hat_s = link.run(n, first_unit=first, draws=draws, want_idx=True)
    assert torch.equal(hat_s, hat) and np.array_equal(c_s, c)
    acc = torch.zeros(4, dtype=torch.int64, device='cuda')       # three ragged shards == one batch
    for a, b in ((0, 5), (5, 6), (6, 32)):
        link.run(b - a, first_unit=a, counters=acc)
    assert np.array_equal(_t(acc), big_c)


def test_full_size_properties():
    """At BASELINE sizes: noiseless frames decode without error; error rate grows with noise;
    repeatable; totals exact."""
    n = 2000
    cfg, siso = make_pair('qam', 64, 1024, 72, 1024, dtype='f32', snr_dB=25.0)
    siso.set_noise_var(0.0)
    c0 = siso.run(n)
    assert c0[2] == n * 1024 and c0[3] == 6 * c0[2] and c0[0] <= 1e-4 * c0[2]
    cfg, link = make_pair('qam', 64, 1024, 72, 1024, Nr=2, Nt=2, dtype='f32', snr_dB=25.0)
    link.set_noise_var(0.0, filter_noise_var=1e-6)
    c0 = link.run(n)
    # without noise only the inter-carrier interference of the time-varying channel is left; the
    # near-ZF filter amplifies it at the rare rank-deficient subcarriers
    assert c0[2] == n * 2048 and c0[3] == 6 * c0[2] and c0[0] <= 3e-4 * c0[2]
    sers = []
    for snr in (30.0, 20.0, 10.0):
        link.set_noise_var(1 / md.dB2Linear(snr))
        c = link.run(n)
        assert np.array_equal(c, link.run(n))                # deterministic
        sers.append(c[0] / c[2])
    assert sers[0] < sers[1] < sers[2] and sers[2] > 0.3


def test_unsupported_shapes_raise():
    from pyphysim_b200 import links
    pm = product_modem('qam', 16)
    with pytest.raises(NotImplementedError):
        links.OfdmTdlLink(pm, 64, 16, 52, Nr=2, Nt=4, tap_powers_linear=[1.0], tap_delays=[0]).run(1)      # Nt > Nr
    with pytest.raises(ValueError):
        links.OfdmTdlLink(pm, 64, 16, 53, tap_powers_linear=[1.0], tap_delays=[0]).run(1)
    with pytest.raises(NotImplementedError):
        links.OfdmTdlLink(pm, 64, 16, 52, tap_powers_linear=[1.0], tap_delays=[0], Fd=1e5, Ts=1e-4,
                          jakes_mode='poly').run(1)
