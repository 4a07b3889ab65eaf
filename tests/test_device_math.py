"""Host-side checks of the hand-rolled device math (no GPU): the constants are parsed out of the CUDA source,
so editing a coefficient without re-deriving it fails here before it reaches a GPU box."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sincospi_kf_coefficients():
    src = open(os.path.join(ROOT, 'pyphysim_b200', 'csrc', 'rng.cuh')).read()
    body = src[src.index('void sincospi_kf('):]
    body = body[:body.index('\n}\n')]
    num = r'(-?\d\.\d+e[+-]\d+)f'
    p = re.search(r'float p = fmaf\(%s, f2, %s\);\s*p = fmaf\(p, f2, %s\);' % (num, num, num), body)
    g = re.search(r'float g = fmaf\(%s, f2, %s\);\s*g = fmaf\(g, f2, %s\);\s*g = fmaf\(g, f2, %s\);' % (num, num, num, num), body)
    pi = re.search(r'f \* (3\.\d+)f', body)
    assert p and g and pi, 'sincospi_kf no longer has the shape this test parses'
    return [np.float32(v) for v in p.groups()], [np.float32(v) for v in g.groups()], np.float32(pi.group(1))


def test_sincospi_kf_polynomials_are_accurate_on_the_reduced_interval():
    """sin(pi f), cos(pi f) for |f| <= 0.2515 (what remains after subtracting k / 2, plus the < 1e-3
    half-turns the ray phases add after the reduction): float32 evaluation within 1.2e-7 absolute
    (the library sincospif is 1-2 ulp = 6e-8..1.2e-7 at values near 1)."""
    (s3, s2, s1), (c4, c3, c2, c1), pi = _sincospi_kf_coefficients()
    f = np.linspace(-0.2515, 0.2515, 1000001).astype(np.float32)
    f2 = f * f
    p = (s3 * f2 + s2) * f2 + s1
    sn = (f2 * f) * p + f * pi
    g = ((c4 * f2 + c3) * f2 + c2) * f2 + c1
    cs = g * f2 + np.float32(1)
    fd = f.astype(np.float64)
    assert np.abs(sn - np.sin(np.pi * fd)).max() < 1.2e-7
    assert np.abs(cs - np.cos(np.pi * fd)).max() < 1.2e-7
    # the approximation error proper (double evaluation of the same float coefficients) is far below that
    pd = (np.float64(s3) * fd ** 2 + np.float64(s2)) * fd ** 2 + np.float64(s1)
    snd = fd ** 3 * pd + fd * np.float64(pi)
    gd = ((np.float64(c4) * fd ** 2 + np.float64(c3)) * fd ** 2 + np.float64(c2)) * fd ** 2 + np.float64(c1)
    csd = gd * fd ** 2 + 1.0
    assert np.abs(snd - np.sin(np.pi * fd)).max() < 3e-8
    assert np.abs(csd - np.cos(np.pi * fd)).max() < 2e-8


def test_sincospi_kf_quadrant_logic():
    """x = k / 2 + f: (sin, cos)(pi x) from (sin, cos)(pi f) by the swap / sign rules the device applies."""
    rng = np.random.default_rng(5)
    k = rng.integers(0, 9, 20000)
    f = rng.uniform(-0.25, 0.25, 20000)
    S, C = np.sin(np.pi * f), np.cos(np.pi * f)
    sw = (k & 1).astype(bool)
    ss, cc = np.where(sw, C, S), np.where(sw, S, C)
    s = np.where(((k.astype(np.uint32) << 30) & 0x80000000) != 0, -ss, ss)
    c = np.where((((k + 1).astype(np.uint32) << 30) & 0x80000000) != 0, -cc, cc)
    x = 0.5 * k + f
    np.testing.assert_allclose(s, np.sin(np.pi * x), atol=1e-12)
    np.testing.assert_allclose(c, np.cos(np.pi * x), atol=1e-12)
