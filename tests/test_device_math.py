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


def _boxmuller_radius_constants():
    src = open(os.path.join(ROOT, 'pyphysim_b200', 'csrc', 'rng.cuh')).read()
    body = src[src.index('float boxmuller_radius(float u)'):]
    body = body[:body.index('\n}\n')]
    num = r'(-?\d\.\d+(?:e[+-]?\d+)?)f'
    magic = re.search(r'\(i - (0x[0-9a-f]+)\)', body)
    init = re.search(r'float r = %s, t = %s;' % (num, num), body)
    steps = re.findall(r'(r|t) = fmaf\((r|t), (s|m), %s\);' % num, body)
    ln2 = re.search(r'fmaf\(float\(e\), %s, r\)' % num, body)
    assert magic and init and len(steps) == 7 and ln2, 'boxmuller_radius no longer has the shape this test parses'
    return int(magic.group(1), 16), [np.float32(v) for v in init.groups()], steps, np.float32(ln2.group(1))


def test_boxmuller_radius_log_is_accurate_to_an_ulp():
    """-ln(u) as rng.cuh::boxmuller_radius evaluates it (exponent split that puts the mantissa in [2/3, 4/3), log1p
    minimax polynomial, e * ln2 / 2^23), over the whole input range u = (x + 0.5) 2^-32 of the generator: every
    float32 step emulated in NumPy, compared with float64 log.  < 1 ulp, and the radius sqrt(-ln u) within one ulp."""
    magic, (r0, t0), steps, ln2s = _boxmuller_radius_constants()
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.integers(0, 2 ** 32, 400000, dtype=np.uint64),
                        np.array([0, 1, 2, 3, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 129, 2 ** 32 - 2, 2 ** 32 - 1], dtype=np.uint64),
                        (2 ** 32 - 1 - rng.integers(0, 2 ** 20, 50000, dtype=np.uint64))])      # u close to 1: -ln u tiny
    u = (x.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)).astype(np.float32)
    assert u.min() >= 2.0 ** -33 and u.max() <= 1.0
    i = u.view(np.int32).astype(np.int64)
    e = (i - magic) & 0xff800000
    e = np.where(e >= 2 ** 31, e - 2 ** 32, e)
    m = ((i - e).astype(np.int32).view(np.float32) - np.float32(1)).astype(np.float32)
    assert m.min() >= -1 / 3 - 1e-6 and m.max() <= 1 / 3 + 1e-6
    s = (m * m).astype(np.float32)
    f32 = lambda v: np.asarray(v, dtype=np.float64).astype(np.float32)       # noqa: E731  (one rounding, like fmaf)
    # the seven constant-addend steps in source order (r, t, r, t, r on s; then r, r on m), with r = fmaf(t, m, r)
    # between them and r = fmaf(r, s, m) at the end
    assert [(d, a, b) for d, a, b, _ in steps] == [('r', 'r', 's'), ('t', 't', 's'), ('r', 'r', 's'), ('t', 't', 's'),
                                                   ('r', 'r', 's'), ('r', 'r', 'm'), ('r', 'r', 'm')]
    r, t = np.full_like(m, r0), np.full_like(m, t0)
    c = [np.float32(v[3]) for v in steps]
    r = f32(r.astype(np.float64) * s + float(c[0])); t = f32(t.astype(np.float64) * s + float(c[1]))
    r = f32(r.astype(np.float64) * s + float(c[2])); t = f32(t.astype(np.float64) * s + float(c[3]))
    r = f32(r.astype(np.float64) * s + float(c[4]))
    r = f32(t.astype(np.float64) * m + r.astype(np.float64))
    r = f32(r.astype(np.float64) * m + float(c[5]))
    r = f32(r.astype(np.float64) * m + float(c[6]))
    r = f32(r.astype(np.float64) * s + m.astype(np.float64))
    nl = -f32(e.astype(np.float64) * float(ln2s) + r.astype(np.float64))
    ref = -np.log(u.astype(np.float64))
    ulp = np.spacing(np.maximum(ref, 1e-30).astype(np.float32)).astype(np.float64)
    assert np.all(np.abs(nl - ref) <= 1.0 * ulp + 1e-45), float((np.abs(nl - ref) / ulp).max())
    rad = np.sqrt(ref)
    assert np.all(np.abs(np.sqrt(np.maximum(nl, 0.0)) - rad) <= np.spacing(rad.astype(np.float32)).astype(np.float64))
    assert abs(float(ln2s) * 2 ** 23 - np.log(2.0)) < 1e-7
