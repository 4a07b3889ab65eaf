import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
w = bench.WORKLOADS['ofdm1024_qam64_mimo2x2_tdl']
link = bench.make_link(w)
n, first = 2000, 77
draws = link.draw(first, n)
F = link.run(n, first_unit=first, want_idx=True, want_eq=True, want_rx=True)
S = link.run(n, first_unit=first, draws=draws, want_idx=True, want_eq=True, want_rx=True)
print(os.environ.get('B200PHY_LIB', 'current'), 'hat diff', int((F[1] != S[1]).sum()), 'eq max diff', float((F[2] - S[2]).abs().max()), 'rx max diff', float((F[3] - S[3]).abs().max()),
      'rx differing', int(((F[3] - S[3]).abs() > 0).sum()), 'of', F[3].numel())
for nv in (1e-30,):
    link.set_noise_var(nv)
    F = link.run(n, first_unit=first, want_idx=True, want_eq=True, want_rx=True)
    S = link.run(n, first_unit=first, draws=draws, want_idx=True, want_eq=True, want_rx=True)
    print('noise_var', nv, 'rx max diff', float((F[3] - S[3]).abs().max()), 'rx differing', int(((F[3] - S[3]).abs() > 0).sum()))
# generic kernel (no pair kernel): fused vs stream
from pyphysim_b200 import links
g = bench.make_link(w); g.params.reserved |= 1
F = g.run(200, first_unit=first, want_idx=True, want_eq=True, want_rx=True)
S = g.run(200, first_unit=first, draws=tuple(d[:200] for d in draws), want_idx=True, want_eq=True, want_rx=True)
print('generic kernel: rx max diff', float((F[3] - S[3]).abs().max()), 'rx differing', int(((F[3] - S[3]).abs() > 0).sum()))
