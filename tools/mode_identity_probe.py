#!/usr/bin/env python
"""Repeat the headline kernel (TC=1: tensor-core H_k variant, TC=0: default) on the same frames: fused vs fused, stream vs
stream (run-to-run determinism) and fused vs stream (the two modes must agree bit for bit): decisions and equalised samples."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
w = bench.WORKLOADS['ofdm1024_qam64_mimo2x2_tdl']
link = bench.make_link(w, tensor_cores=(os.environ.get('TC', '1') == '1'))
n, first = 6000, 77
draws = link.draw(first, n)
F = [link.run(n, first_unit=first, want_idx=True, want_eq=True) for _ in range(3)]
S = [link.run(n, first_unit=first, draws=draws, want_idx=True, want_eq=True) for _ in range(3)]
def diff(a, b):
    return int((a[1] != b[1]).sum()), float((a[2] - b[2]).abs().max())
print('fused  vs fused :', [diff(F[0], F[i]) for i in (1, 2)])
print('stream vs stream:', [diff(S[0], S[i]) for i in (1, 2)])
print('fused  vs stream:', [diff(F[i], S[i]) for i in range(3)])
d = (F[0][2] - S[0][2]).abs()
bad = torch.nonzero(d.reshape(n, -1).max(dim=1).values > 0).reshape(-1)
print('frames with differing equalised symbols:', bad[:20].tolist(), 'of', n, 'count', bad.numel())
if bad.numel():
    fr = int(bad[0]); row = d.reshape(n, -1)[fr]
    idx = torch.nonzero(row > 0).reshape(-1)
    print('frame', fr, 'differing symbols', idx.numel(), 'first', idx[:16].tolist(), 'max', float(row.max()))
