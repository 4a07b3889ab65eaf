#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

It imports ``pyphysim`` from /root/reference, feeds its classes deterministic
inputs derived from the shared Philox stream (oracle/philox.py) and stores the
reference's outputs as small ``.npz`` files.  The fixtures travel with the repo;
the reference does not.  tests/test_oracle_golden.py checks the oracle against
them, the GPU parity tests check the CUDA path against the oracle and against
these files.
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')

from pyphysim.channels import fading, fading_generators  # noqa: E402
from pyphysim.mimo import mimo as rmimo  # noqa: E402
from pyphysim.modulators import fundamental, ofdm as rofdm  # noqa: E402
from pyphysim.util import conversion, misc  # noqa: E402

from oracle import philox  # noqa: E402

SEED = 0xC0FFEE


class QueueRS:
    """Stand-in for np.random.RandomState: ``rand(*shape)`` serves the queued
    arrays in order (JakesSampleGenerator accepts an injected RS,
    channels/fading_generators.py:335-341)."""

    def __init__(self):
        self.queue = []

    def rand(self, *shape):
        n = int(np.prod(shape))
        if self.queue and self.queue[0].size == n:
            return self.queue.pop(0).reshape(shape)
        return np.full(shape, 0.25)          # constructor's throw-away draw


def uniforms(unit, shape):
    """phi/(2 pi), psi/(2 pi) exactly as oracle.philox.jakes_phases lays them out."""
    phi, psi = philox.jakes_phases(SEED, [unit], shape)
    return phi[0] / (2 * np.pi), psi[0] / (2 * np.pi)


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('%-22s %7.1f KiB' % (name + '.npz', os.path.getsize(path) / 1024))


def gen_constellations():
    out = {}
    for M in (4, 16, 64, 256):
        out['qam%d' % M] = fundamental.QAM(M).symbols
    for M in (2, 4, 8, 16):
        out['psk%d' % M] = fundamental.PSK(M).symbols
    out['psk8_off'] = fundamental.PSK(8, 0.3).symbols
    p = fundamental.PSK(8)
    p.setPhaseOffset(0.2)                      # drops the Gray order (fundamental.py:459)
    out['psk8_setoffset'] = p.symbols
    out['qpsk'] = fundamental.QPSK().symbols
    out['bpsk'] = fundamental.BPSK().symbols
    out['gray16'] = conversion.binary2gray(np.arange(16))
    out['ungray16'] = conversion.gray2binary(np.arange(16))
    save('constellations', **out)


def gen_demap():
    out = {}
    mods = {'qam16': fundamental.QAM(16), 'qam64': fundamental.QAM(64),
            'qam256': fundamental.QAM(256), 'psk8': fundamental.PSK(8),
            'qpsk': fundamental.QPSK(), 'bpsk': fundamental.BPSK()}
    for i, (name, mod) in enumerate(mods.items()):
        n = 4096
        idx = philox.data_indices(SEED, [100 + i], n, misc.level2bits(mod.M))[0]
        noise = philox.cnormal(SEED, 2, [100 + i], n)[0]
        r = mod.modulate(idx) + 0.35 * noise
        out[name + '_idx'] = idx
        out[name + '_r'] = r
        out[name + '_hat'] = mod.demodulate(r)
        out[name + '_biterr'] = np.array(misc.count_bit_errors(idx, out[name + '_hat']))
    a = philox.words(SEED, 0, [7], 64)[0].astype(np.int64)
    out['count_bits_in'] = a
    out['count_bits_out'] = misc.count_bits(a)
    save('demap', **out)


def gen_ofdm():
    out = {}
    for tag, (f, c, u) in {'a': (64, 16, 52), 'b': (64, 4, 64), 'c': (1024, 72, 1024),
                           'd': (128, 0, 60)}.items():
        o = rofdm.OFDM(f, c, u)
        x = philox.cnormal(SEED, 1, [200], 3 * u - 5)[0]        # needs zero padding
        out[tag + '_params'] = np.array([f, c, u])
        out[tag + '_bins'] = o.get_used_subcarrier_indexes()
        out[tag + '_x'] = x
        out[tag + '_mod'] = o.modulate(x)
        r = philox.cnormal(SEED, 2, [201], 3 * (f + c))[0]
        out[tag + '_r'] = r
        out[tag + '_demod'] = o.demodulate(r.copy())
    save('ofdm', **out)


def gen_fading():
    out = {}
    # profile discretisation
    for pname, prof in (('tu', fading.COST259_TUx), ('ra', fading.COST259_RAx),
                        ('ht', fading.COST259_HTx)):
        for tname, Ts in (('2048', 1 / (15e3 * 2048)), ('1024', 1 / (15e3 * 1024)),
                          ('128', 1 / (15e3 * 128))):
            d = prof.get_discretize_profile(Ts)
            out['%s_%s_powers' % (pname, tname)] = d.tap_powers_linear
            out['%s_%s_delays' % (pname, tname)] = d.tap_delays
    # Jakes generator alone, shape None, two consecutive calls + a skip
    rs = QueueRS()
    u_phi, u_psi = uniforms(300, (8, 1))
    rs.queue = [u_phi, u_psi]
    g = fading_generators.JakesSampleGenerator(Fd=100.0, Ts=1e-3, L=8, RS=rs)
    out['j0_phi'], out['j0_psi'] = 2 * np.pi * u_phi[:, 0], 2 * np.pi * u_psi[:, 0]
    g.generate_more_samples(50)
    out['j0_h1'] = g.get_samples()
    g.skip_samples_for_next_generation(7)
    g.generate_more_samples(20)
    out['j0_h2'] = g.get_samples()
    out['j0_t_end'] = np.array(g._current_time)
    # Jakes with a MIMO shape
    rs = QueueRS()
    g = fading_generators.JakesSampleGenerator(Fd=30.0, Ts=5e-6, L=20, RS=rs)
    u_phi, u_psi = uniforms(301, (20, 3, 2, 1))
    rs.queue = [u_phi, u_psi]
    g.shape = (3, 2)
    g.generate_more_samples(40)
    out['j1_phi'], out['j1_psi'] = 2 * np.pi * u_phi[..., 0], 2 * np.pi * u_psi[..., 0]
    out['j1_h'] = g.get_samples()
    save('fading', **out)


def gen_tdl():
    out = {}
    # SISO: QAM64 + OFDM(128,16,100) x 3 symbols over TU @ Ts=1/(15e3*128), equalised
    fft, cp, used, nsym = 128, 16, 100, 3
    Ts = 1 / (15e3 * fft)
    qam = fundamental.QAM(64)
    o = rofdm.OFDM(fft, cp, used)
    idx = philox.data_indices(SEED, [400], nsym * used, 6)[0]
    tx = o.modulate(qam.modulate(idx))
    rs = QueueRS()
    jakes = fading_generators.JakesSampleGenerator(Fd=200.0, Ts=Ts, L=20, RS=rs)
    prof = fading.COST259_TUx.get_discretize_profile(Ts)
    u_phi, u_psi = uniforms(400, (20, prof.num_taps, 1))
    rs.queue = [u_phi, u_psi]
    ch = fading.TdlChannel(jakes, prof)
    rx = ch.corrupt_data(tx)
    noise = philox.noise_rows(SEED, [400], 1, rx.size)[0, 0]
    nv = 1e-3
    rxn = rx + math.sqrt(nv) * noise
    Y = o.demodulate(rxn[:tx.size].copy())
    ir = ch.get_last_impulse_response()
    eq = rofdm.OfdmOneTapEqualizer(o).equalize_data(Y, ir)
    out.update(s_params=np.array([fft, cp, used, nsym]), s_Ts=np.array(Ts), s_Fd=np.array(200.0),
               s_idx=idx, s_tx=tx, s_phi=2 * np.pi * u_phi[..., 0], s_psi=2 * np.pi * u_psi[..., 0],
               s_taps=ir.tap_values_sparse, s_rx=rx, s_noise=noise, s_nv=np.array(nv), s_Y=Y,
               s_eq=eq, s_hat=qam.demodulate(eq), s_delays=prof.tap_delays,
               s_powers=prof.tap_powers_linear)
    # MIMO 3x2 corrupt_data (the shape the reference's own test uses), RA profile
    Ts = 1 / (15e3 * 2048)
    rs = QueueRS()
    jakes = fading_generators.JakesSampleGenerator(Fd=50.0, Ts=Ts, L=16, shape=(3, 2), RS=rs)
    prof = fading.COST259_RAx.get_discretize_profile(Ts)
    u_phi, u_psi = uniforms(401, (16, prof.num_taps, 3, 2, 1))
    rs.queue = [u_phi, u_psi]
    ch = fading.TdlMimoChannel(jakes, prof)
    x = philox.cnormal(SEED, 1, [402], 2 * 150)[0].reshape(2, 150)
    y = ch.corrupt_data(x)
    ir = ch.get_last_impulse_response()
    out.update(m_Ts=np.array(Ts), m_Fd=np.array(50.0), m_x=x, m_y=y,
               m_phi=2 * np.pi * u_phi[..., 0], m_psi=2 * np.pi * u_psi[..., 0],
               m_taps=ir.tap_values_sparse, m_delays=prof.tap_delays,
               m_powers=prof.tap_powers_linear,
               m_freq=ir.get_freq_response(64)[:, :, :, ::50])
    save('tdl', **out)


def gen_mimo():
    out = {}
    H43 = philox.cnormal(SEED, 1, [500], 12)[0].reshape(4, 3)
    H44 = philox.cnormal(SEED, 1, [501], 16)[0].reshape(4, 4)
    H22 = philox.cnormal(SEED, 1, [502], 4)[0].reshape(2, 2)
    H32 = philox.cnormal(SEED, 1, [503], 6)[0].reshape(3, 2)
    H12 = philox.cnormal(SEED, 1, [504], 2)[0].reshape(1, 2)
    for name, H in (('h43', H43), ('h44', H44), ('h22', H22)):
        Nr, Nt = H.shape
        b = rmimo.Blast(H)
        s = philox.cnormal(SEED, 0, [510], 5 * Nt)[0]
        x = b.encode(s)
        y = H @ x + 0.05 * philox.cnormal(SEED, 2, [511], Nr * 5)[0].reshape(Nr, 5)
        out[name] = H
        out[name + '_s'] = s
        out[name + '_x'] = x
        out[name + '_y'] = y
        out[name + '_zf'] = b.decode(y)
        b.set_noise_var(0.01)
        out[name + '_mmse'] = b.decode(y)
    for name, H in (('a22', H22), ('a32', H32), ('a12', H12)):
        Nr = H.shape[0]
        a = rmimo.Alamouti(H)
        s = philox.cnormal(SEED, 0, [520], 8)[0]
        x = a.encode(s)
        y = H @ x + 0.05 * philox.cnormal(SEED, 2, [521], Nr * 8)[0].reshape(Nr, 8)
        out[name] = H
        out[name + '_s'] = s
        out[name + '_x'] = x
        out[name + '_y'] = y
        out[name + '_dec'] = a.decode(y)
    save('mimo', **out)


def gen_links():
    """Whole links through the reference's classes on Philox draws (few units)."""
    from oracle import links as L      # only for the draw layout + configs
    out = {}
    # C3 at full size: QAM64, OFDM(1024,72,1024), TU, Jakes(10 Hz, L=20), 20 dB, 2 frames
    m = L.Modem('qam', 64)
    cfg = L.OfdmTdlConfig(m, 1024, 72, 1024, n_sym=1, noise_var=1 / conversion.dB2Linear(20.0))
    units = np.array([0, 1])
    idx, phi, psi, noise = L.draws_ofdm_tdl(cfg, SEED, units)
    qam = fundamental.QAM(64)
    o = rofdm.OFDM(1024, 72, 1024)
    prof = fading.COST259_TUx.get_discretize_profile(cfg.Ts)
    hats, eqs = [], []
    for u in range(2):
        rs = QueueRS()
        jakes = fading_generators.JakesSampleGenerator(cfg.Fd, cfg.Ts, cfg.L, RS=rs)
        rs.queue = [phi[u][..., None] / (2 * np.pi), psi[u][..., None] / (2 * np.pi)]
        ch = fading.TdlChannel(jakes, prof)
        tx = o.modulate(qam.modulate(idx[u]))
        rx = ch.corrupt_data(tx)
        rx += math.sqrt(cfg.noise_var) * noise[u, 0]
        Y = o.demodulate(rx[0:tx.size].copy())
        eq = rofdm.OfdmOneTapEqualizer(o).equalize_data(Y, ch.get_last_impulse_response())
        eqs.append(eq)
        hats.append(qam.demodulate(eq))
    out.update(c3_units=units, c3_idx=idx, c3_eq=np.array(eqs), c3_hat=np.array(hats))

    # 2x2 Blast-MMSE + OFDM(256,18,200) x 2 symbols over TU, QAM16, 25 dB, 2 frames
    m = L.Modem('qam', 16)
    cfg = L.OfdmTdlConfig(m, 256, 18, 200, n_sym=2, Nr=2, Nt=2, Fd=300.0,
                          noise_var=1 / conversion.dB2Linear(25.0))
    units = np.array([10, 11])
    idx, phi, psi, noise = L.draws_ofdm_tdl(cfg, SEED, units)
    qam = fundamental.QAM(16)
    o = rofdm.OFDM(256, 18, 200)
    prof = fading.COST259_TUx.get_discretize_profile(cfg.Ts)
    bins = o.get_used_subcarrier_indexes()
    hats, eqs = [], []
    for u in range(2):
        rs = QueueRS()
        jakes = fading_generators.JakesSampleGenerator(cfg.Fd, cfg.Ts, cfg.L, shape=(2, 2), RS=rs)
        rs.queue = [phi[u][..., None] / (2 * np.pi), psi[u][..., None] / (2 * np.pi)]
        ch = fading.TdlMimoChannel(jakes, prof)
        blast = rmimo.Blast()
        blast.set_channel_matrix(np.eye(2, dtype=complex))
        layers = blast.encode(qam.modulate(idx[u]))
        tx = np.stack([o.modulate(layers[t]) for t in range(2)])
        rx = ch.corrupt_data(tx)
        rx += math.sqrt(cfg.noise_var) * noise[u]
        Y = np.stack([o.demodulate(rx[r, 0:tx.shape[1]].copy()) for r in range(2)])
        ir = ch.get_last_impulse_response()
        Hf = ir.get_freq_response(256)                        # [fft, 2, 2, N]
        Hf = Hf.reshape(256, 2, 2, cfg.n_sym, -1).mean(axis=-1)
        eq = np.empty(cfg.n_data, dtype=complex)
        blast.set_noise_var(cfg.noise_var)
        for sy in range(cfg.n_sym):
            for q in range(200):
                blast.set_channel_matrix(Hf[bins[q], :, :, sy])
                j = sy * 200 + q
                eq[2 * j:2 * j + 2] = blast.decode(Y[:, j].reshape(2, 1))
        eqs.append(eq)
        hats.append(qam.demodulate(eq))
    out.update(m2_units=units, m2_idx=idx, m2_eq=np.array(eqs), m2_hat=np.array(hats))

    # Alamouti 2x2 QPSK, 10 dB, 64 realizations of one codeword (C4 shape)
    units = np.arange(64)
    idx, H, n = L.draws_flat_mimo(SEED, units, 2, 2, 2, 2, 2)
    qpsk = fundamental.QPSK()
    nv = 1 / conversion.dB2Linear(10.0)
    hats = []
    for u in range(64):
        a = rmimo.Alamouti()
        a.set_channel_matrix(H[u])
        y = np.dot(H[u], a.encode(qpsk.modulate(idx[u]))) + n[u] * np.sqrt(nv)
        hats.append(qpsk.demodulate(a.decode(y)))
    out.update(c4_idx=idx, c4_hat=np.array(hats))

    # 64-QAM flat Rayleigh, 10 dB, 4096 realizations (C2 shape)
    units = np.arange(4096)
    idx, h, n = L.draws_siso_flat(SEED, units, 6)
    qam = fundamental.QAM(64)
    r = h * qam.modulate(idx) + math.sqrt(1 / conversion.dB2Linear(10.0)) * n
    r /= h
    out.update(c2_idx=idx, c2_hat=qam.demodulate(r))
    save('links', **out)


if __name__ == '__main__':
    gen_constellations()
    gen_demap()
    gen_ofdm()
    gen_fading()
    gen_tdl()
    gen_mimo()
    gen_links()
