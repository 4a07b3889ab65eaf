#!/usr/bin/env python
"""Golden fixture for TdlChannel.corrupt_data_in_freq_domain, produced by the
unmodified reference in the build container:  python tests/golden/make_golden_freqdomain.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')

from pyphysim.channels import fading, fading_generators  # noqa: E402

from make_golden import SEED, QueueRS, uniforms  # noqa: E402
from oracle import philox  # noqa: E402

out = {}
# SISO, TU profile, 64-point blocks on a carrier subset, 5 blocks
Ts = 1 / (15e3 * 1024)
rs = QueueRS()
jakes = fading_generators.JakesSampleGenerator(Fd=80.0, Ts=Ts, L=20, RS=rs)
prof = fading.COST259_TUx.get_discretize_profile(Ts)
u_phi, u_psi = uniforms(600, (20, prof.num_taps, 1))
rs.queue = [u_phi, u_psi]
ch = fading.TdlChannel(jakes, prof)
car = np.r_[1:25, 40:64]
x = philox.cnormal(SEED, 1, [601], 5 * car.size)[0]
y = ch.corrupt_data_in_freq_domain(x, 64, car)
out.update(s_Ts=np.array(Ts), s_phi=2 * np.pi * u_phi[..., 0], s_psi=2 * np.pi * u_psi[..., 0], s_x=x, s_y=y,
           s_car=car, s_taps=ch.get_last_impulse_response().tap_values_sparse, s_t_end=np.array(jakes._current_time))
y2 = ch.corrupt_data_in_freq_domain(x[:128], 64)                  # continues the generator's clock, full blocks
out.update(s_y2=y2, s_taps2=ch.get_last_impulse_response().tap_values_sparse)
# MIMO 3x2, RA profile, slice of carriers
Ts = 1 / (15e3 * 2048)
rs = QueueRS()
jakes = fading_generators.JakesSampleGenerator(Fd=120.0, Ts=Ts, L=16, shape=(3, 2), RS=rs)
prof = fading.COST259_RAx.get_discretize_profile(Ts)
u_phi, u_psi = uniforms(602, (16, prof.num_taps, 3, 2, 1))
rs.queue = [u_phi, u_psi]
ch = fading.TdlMimoChannel(jakes, prof)
x = philox.cnormal(SEED, 1, [603], 2 * 4 * 32)[0].reshape(2, 128)
y = ch.corrupt_data_in_freq_domain(x, 128, slice(10, 42))
out.update(m_Ts=np.array(Ts), m_phi=2 * np.pi * u_phi[..., 0], m_psi=2 * np.pi * u_psi[..., 0], m_x=x, m_y=y,
           m_taps=ch.get_last_impulse_response().tap_values_sparse)
ch.switched_direction = True
x3 = philox.cnormal(SEED, 1, [604], 3 * 64)[0].reshape(3, 64)
out.update(m_x3=x3, m_y3=ch.corrupt_data_in_freq_domain(x3, 128, slice(10, 42)))
path = os.path.join(HERE, 'freqdomain.npz')
np.savez_compressed(path, **out)
print('freqdomain.npz %.1f KiB' % (os.path.getsize(path) / 1024))
