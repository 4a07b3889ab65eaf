"""Helpers shared by the GPU parity tests."""
import numpy as np

from oracle import links as OL
from oracle import philox


def product_modem(kind, M=2, phase_offset=0.0):
    from pyphysim_b200.modulators import fundamental as F
    if kind == 'qam':
        return F.QAM(M)
    if kind == 'bpsk':
        return F.BPSK()
    if kind == 'qpsk':
        return F.QPSK()
    return F.PSK(M, phase_offset)


def oracle_modem(kind, M=2, phase_offset=0.0):
    if kind == 'qpsk':
        return OL.Modem('psk', 4, np.pi / 4)
    return OL.Modem(kind, M, phase_offset)


def cuda(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def assert_samples_close(dev, ref, rel=1e-5, what=''):
    """|dev - ref| <= rel * max(|ref|, rms(ref)) (SURVEY.md §7 hard part 2: pointwise relative error
    is meaningless in deep fades, so the floor is the rms)."""
    dev = np.asarray(dev).astype(np.complex128).reshape(-1)
    ref = np.asarray(ref).astype(np.complex128).reshape(-1)
    rms = np.sqrt(np.mean(np.abs(ref) ** 2))
    tol = rel * np.maximum(np.abs(ref), rms)
    err = np.abs(dev - ref)
    worst = np.argmax(err / tol)
    assert np.all(err <= tol), '%s: worst err %.3g vs tol %.3g at %d' % (what, err[worst], tol[worst], worst)


def decision_margin(modem, r):
    """Distance gap between the best and second-best constellation points (oracle side)."""
    if modem.kind == 'bpsk':
        return 2 * np.abs(np.real(r))
    d = np.abs(modem.symbols.reshape(-1, 1) - np.asarray(r).reshape(1, -1))
    d.sort(axis=0)
    return d[1] - d[0]


def assert_decisions(dev_idx, ref_idx, modem, ref_samples, exact, eps=2e-4, what=''):
    """exact: indices identical.  Otherwise (float32 arithmetic vs the float64 oracle) any mismatch
    must sit on a decision boundary: oracle margin below eps, and be rare."""
    dev_idx = np.asarray(dev_idx).reshape(-1).astype(np.int64)
    ref_idx = np.asarray(ref_idx).reshape(-1).astype(np.int64)
    bad = np.nonzero(dev_idx != ref_idx)[0]
    if exact:
        assert bad.size == 0, '%s: %d index mismatches (first at %s)' % (what, bad.size, bad[:5])
        return 0
    if bad.size:
        marg = decision_margin(modem, np.asarray(ref_samples).reshape(-1)[bad])
        assert np.all(marg < eps), '%s: mismatch with margin %.3g' % (what, marg.max())
        assert bad.size <= max(2, 1e-3 * dev_idx.size), '%s: %d mismatches' % (what, bad.size)
    return bad.size


def host_draws(kind, seed, units, **kw):
    return getattr(OL, 'draws_' + kind)(seed=seed, units=units, **kw)


__all__ = ['philox']
