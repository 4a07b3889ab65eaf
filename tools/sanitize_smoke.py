#!/usr/bin/env python
"""Small invocations of every fused link (both OFDM kernels, both dtypes, fused + stream) to be run under
compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyphysim_b200 import links                       # noqa: E402
from pyphysim_b200.channels.fading import COST259_TUx  # noqa: E402
from pyphysim_b200.modulators import QAM, QPSK         # noqa: E402


def ofdm(fft, cp, used, Nr, Nt, nsym, dtype, pair=True, jakes='auto', Fd=10.0):
    Ts = 1.0 / (15e3 * fft)
    prof = COST259_TUx.get_discretize_profile(Ts)
    return links.OfdmTdlLink(QAM(16), fft, cp, used, num_ofdm_symbols=nsym, Nr=Nr, Nt=Nt,
                             tap_powers_linear=prof.tap_powers_linear, tap_delays=prof.tap_delays, Fd=Fd, Ts=Ts,
                             L=20, noise_var=0.01, dtype=dtype, jakes_mode=jakes, use_pair_kernel=pair)


cases = [ofdm(1024, 72, 1024, 2, 2, 2, 'f32'), ofdm(1024, 72, 1024, 2, 2, 1, 'f32', pair=False),
         ofdm(1024, 72, 1024, 2, 2, 1, 'f32'),          # headline shape: TMA bulk input pipeline in stream mode
         ofdm(1024, 72, 600, 1, 1, 2, 'f32'), ofdm(256, 18, 200, 2, 2, 2, 'f64'),
         ofdm(256, 18, 200, 2, 1, 1, 'f32', jakes='recurrence', Fd=500.0), ofdm(2048, 144, 2048, 4, 4, 1, 'f32'),
         ofdm(512, 36, 300, 4, 2, 1, 'f64'),
         ofdm(1024, 72, 1024, 1, 1, 1, 'f32'),          # SISO frame-pair kernel (odd n: last frame generic)
         ofdm(1024, 72, 1024, 4, 2, 1, 'f32')]          # 4x2 pair kernel
for link in cases:
    n = 5
    c_f = link.run(n, first_unit=3)
    d = link.draw(3, n)
    c_s = link.run(n, first_unit=3, draws=d, want_idx=True)[0]
    assert np.array_equal(c_f, c_s), (c_f, c_s)
print('ofdm ok')
for dt in ('f32', 'f64'):
    assert links.link_siso_flat(QAM(64), 0.05, 1003, dtype=dt)[2] == 1003
    assert links.link_alamouti(QPSK(), 0.1, 777, Nr=2, num_symbols=4, dtype=dt)[2] == 777 * 4
    assert links.link_blast(QAM(16), 0.05, 555, Nr=4, Nt=3, num_symbols=2, filter_noise_var=0.05, dtype=dt)[2] == 555 * 6
for sch in ('svd', 'gmd', 'mrt'):
    kw = dict(scheme=sch, Nr=(1 if sch == 'mrt' else 3), Nt=3, num_symbols=2)
    c = links.link_precoded(QAM(16), 0.05, 333, dtype='f32', **kw)
    assert c[2] > 0
# reference signals + estimators (stage kernels of row next-4)
from pyphysim_b200.reference_signals.channel_estimation import CazacBasedChannelEstimator   # noqa: E402
from pyphysim_b200.reference_signals.root_sequence import RootSequence                       # noqa: E402
from pyphysim_b200.reference_signals.srs import SrsUeSequence                                # noqa: E402
u = SrsUeSequence(RootSequence(11, 300), 5, normalize=True)
Y = (torch.randn(7, 2, 300, dtype=torch.float64) + 1j * torch.randn(7, 2, 300, dtype=torch.float64)).cuda()
H = CazacBasedChannelEstimator(u).estimate_batch(Y.to(torch.complex64), 20)
assert tuple(H.shape) == (7, 2, 600)
torch.cuda.synchronize()
print('sanitize smoke ok')
