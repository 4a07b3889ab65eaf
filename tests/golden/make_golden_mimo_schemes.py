#!/usr/bin/env python
"""Golden fixture for MRT / MRC / SVDMimo / GMDMimo, util.misc.gmd and calc_post_processing_SINRs
(SURVEY.md §8f next-3), produced by the unmodified reference in the build container:
    python tests/golden/make_golden_mimo_schemes.py
Channels and data come from the oracle's Philox streams so every case can be regenerated from its seed."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')

from pyphysim.mimo import mimo  # noqa: E402
from pyphysim.util import misc  # noqa: E402

from make_golden import SEED  # noqa: E402
from oracle import philox  # noqa: E402

out = {}
NV = 0.02


def chan(unit, Nr, Nt):
    return philox.cnormal(SEED, 1, [unit], Nr * Nt)[0].reshape(Nr, Nt)


def data(unit, n):
    return philox.cnormal(SEED, 0, [unit], n)[0]


# square channels: SVD and GMD
for k, n in enumerate((2, 3, 4)):
    H = chan(700 + k, n, n)
    x = data(710 + k, 5 * n)
    noise = np.sqrt(NV) * philox.cnormal(SEED, 2, [720 + k], 5 * n)[0].reshape(n, 5)
    U, S, Vh = np.linalg.svd(H)
    Q, R, P = misc.gmd(U, S, Vh)
    pre = 'sq%d_' % n
    out.update({pre + 'H': H, pre + 'x': x, pre + 'noise': noise, pre + 'U': U, pre + 'S': S, pre + 'Vh': Vh,
                pre + 'Q': Q, pre + 'R': R, pre + 'P': P})
    for name, cls in (('svd', mimo.SVDMimo), ('gmd', mimo.GMDMimo)):
        obj = cls(H)
        obj.set_noise_var(NV)
        W = cls._calc_precoder(H)
        G = cls._calc_receive_filter(H, NV)
        enc = obj.encode(x)
        dec = obj.decode(H.dot(enc) + noise)
        out.update({pre + name + '_W': W, pre + name + '_G': G, pre + name + '_enc': enc, pre + name + '_dec': dec,
                    pre + name + '_sinr_lin': obj.calc_linear_SINRs(NV), pre + name + '_sinr_dB': obj.calc_SINRs(NV),
                    pre + name + '_layers': np.array(obj.getNumberOfLayers())})

# tall channel through gmd() alone (the MIMO classes are only exercised square by the reference's tests)
H = chan(730, 4, 2)
U, S, Vh = np.linalg.svd(H)
Q, R, P = misc.gmd(U, S, Vh)
out.update(tall_H=H, tall_U=U, tall_S=S, tall_Vh=Vh, tall_Q=Q, tall_R=R, tall_P=P)
# a tolerance that drops the smallest singular value
H = chan(731, 3, 3)
U, S, Vh = np.linalg.svd(H)
tol = 0.5 * (S[1] + S[2])
Q, R, P = misc.gmd(U, S, Vh, tol)
out.update(tol_U=U, tol_S=S, tol_Vh=Vh, tol_tol=np.array(tol), tol_Q=Q, tol_R=R, tol_P=P)

# MRT: single receive antenna, 1-D and [1, Nt] channels
for k, nt in enumerate((2, 3, 4)):
    h = chan(740 + k, 1, nt)
    x = data(750 + k, 7)
    noise = np.sqrt(NV) * philox.cnormal(SEED, 2, [760 + k], 7)[0]
    obj = mimo.MRT(h[0] if k == 0 else h)
    enc = obj.encode(x)
    dec = obj.decode(h.dot(enc) + noise)
    pre = 'mrt%d_' % nt
    out.update({pre + 'h': h, pre + 'x': x, pre + 'noise': noise, pre + 'W': mimo.MRT._calc_precoder(h),
                pre + 'G': np.array(mimo.MRT._calc_receive_filter(h)), pre + 'enc': enc, pre + 'dec': dec,
                pre + 'sinr_lin': obj.calc_linear_SINRs(NV)})

# MRC: Blast with a column (1-D) channel and with a tall matrix
h = chan(770, 4, 1)[:, 0]
x = data(771, 6)
noise = np.sqrt(NV) * philox.cnormal(SEED, 2, [772], 24)[0].reshape(4, 6)
obj = mimo.MRC(h)
obj.set_noise_var(NV)
enc = obj.encode(x)
out.update(mrc_h=h, mrc_x=x, mrc_noise=noise, mrc_enc=enc, mrc_dec=obj.decode(h[:, None].dot(enc) + noise),
           mrc_sinr_lin=obj.calc_linear_SINRs(NV), mrc_layers=np.array(obj.getNumberOfLayers()))
H = chan(773, 4, 3)
x = data(774, 9)
noise = np.sqrt(NV) * philox.cnormal(SEED, 2, [775], 12)[0].reshape(4, 3)
obj = mimo.MRC(H)
obj.set_noise_var(None)                       # zero-forcing filter
enc = obj.encode(x)
out.update(mrc2_H=H, mrc2_x=x, mrc2_noise=noise, mrc2_enc=enc, mrc2_dec=obj.decode(H.dot(enc) + noise))
out['noise_var'] = np.array(NV)

np.savez_compressed(os.path.join(HERE, 'mimo_schemes.npz'), **out)
print('wrote mimo_schemes.npz with', len(out), 'arrays')
