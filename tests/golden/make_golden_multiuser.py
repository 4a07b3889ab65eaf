#!/usr/bin/env python
"""Golden fixture for MuChannel / MuMimoChannel (multiuser.py:42-586), produced by the
unmodified reference in the build container:  python tests/golden/make_golden_multiuser.py
Every link's Jakes phases are overwritten (after construction, when the generator shapes are final) with
the oracle's Philox draws for unit `base + rx * num_tx + tx`, so the product can be given the same ones."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')

from pyphysim.channels import fading, fading_generators, multiuser  # noqa: E402

from make_golden import SEED  # noqa: E402
from oracle import philox  # noqa: E402


def set_phases(mu, base):
    """-> phi, psi arrays [num_rx, num_tx, ...generator phase shape]"""
    num_rx, num_tx = mu._su_siso_channels.shape
    phis, psis = [], []
    for rx in range(num_rx):
        for tx in range(num_tx):
            gen = mu._su_siso_channels[rx, tx]._tdlchannel._fading_generator
            shape = gen._phi_l.shape
            phi, psi = philox.jakes_phases(SEED, [base + rx * num_tx + tx], shape[:-1])
            gen._phi_l = phi[0].reshape(shape)
            gen._psi_l = psi[0].reshape(shape)
            phis.append(gen._phi_l.copy())
            psis.append(gen._psi_l.copy())
    return np.array(phis), np.array(psis)


def sig(unit, rows, n):
    return philox.cnormal(SEED, 0, [unit], rows * n)[0].reshape(rows, n)


out = {}
Ts = 1 / (15e3 * 1024)
# ---- 2 receivers x 3 transmitters, SISO links, TU profile, time domain then frequency domain
jakes = fading_generators.JakesSampleGenerator(Fd=70.0, Ts=Ts, L=8)
mu = multiuser.MuChannel((2, 3), jakes, channel_profile=fading.COST259_TUx, Ts=Ts)
phi, psi = set_phases(mu, 900)
pl = np.array([[1.0, 0.3, 0.05], [0.2, 0.8, 0.6]])
mu.set_pathloss(pl)
x = sig(910, 3, 150)
y = mu.corrupt_data(x)
out.update(a_Ts=np.array(Ts), a_phi=phi, a_psi=psi, a_pl=pl, a_x=x, a_y0=y[0], a_y1=y[1],
           a_ir12=mu.get_last_impulse_response(1, 2).tap_values_sparse,
           a_num_taps=np.array(mu.num_taps), a_pad=np.array(mu.num_taps_with_padding))
x2 = sig(911, 3, 3 * 64)
y2 = mu.corrupt_data_in_freq_domain(x2, 64)
car = np.r_[2:30]
x3 = sig(912, 3, 2 * car.size)
y3 = mu.corrupt_data_in_freq_domain(x3, 64, car)
out.update(a_x2=x2, a_y2_0=y2[0], a_y2_1=y2[1], a_car=car, a_x3=x3, a_y3_0=y3[0], a_y3_1=y3[1])
mu.switched_direction = True                       # 3 "receivers" (the transmitters) x 2
x4 = sig(913, 2, 40)
y4 = mu.corrupt_data(x4)
out.update(a_x4=x4, a_y4_0=y4[0], a_y4_1=y4[1], a_y4_2=y4[2])

# ---- 2 x 2 users with 2 x 3 MIMO links, flat-ish 2-tap profile given as tap powers / delays
jakes = fading_generators.JakesSampleGenerator(Fd=30.0, Ts=Ts, L=6)
mm = multiuser.MuMimoChannel(2, 2, 3, jakes, tap_powers_dB=np.array([0.0, -6.0]), tap_delays=np.array([0.0, 3 * Ts]), Ts=Ts)
phi, psi = set_phases(mm, 950)
xs = np.empty(2, dtype=object)
xs[0] = sig(960, 3, 60)
xs[1] = sig(961, 3, 60)
ym = mm.corrupt_data(xs)
out.update(b_phi=phi, b_psi=psi, b_x0=xs[0], b_x1=xs[1], b_y0=ym[0], b_y1=ym[1],
           b_ntx=mm.num_tx_antennas, b_nrx=mm.num_rx_antennas,
           b_ir01=mm.get_last_impulse_response(0, 1).tap_values_sparse)

# ---- default construction: flat Rayleigh links cannot be seeded; only shapes are recorded
m0 = multiuser.MuChannel(3)
y0 = m0.corrupt_data(sig(970, 3, 10))
out.update(c_len=np.array([v.size for v in y0]))

np.savez_compressed(os.path.join(HERE, 'multiuser.npz'), **out)
print('wrote multiuser.npz with', len(out), 'arrays,', os.path.getsize(os.path.join(HERE, 'multiuser.npz')) // 1024, 'KiB')
