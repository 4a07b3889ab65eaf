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


def assert_samples_close(dev, ref, rel=1e-5, what='', tol=None):
    """|dev - ref| <= rel * max(|ref|, rms(ref)) (SURVEY.md §7 hard part 2: pointwise relative error
    is meaningless in deep fades, so the floor is the rms).  `tol`: explicit per-sample bound instead.
    Returns the worst err / tol ratio."""
    dev = np.asarray(dev).astype(np.complex128).reshape(-1)
    ref = np.asarray(ref).astype(np.complex128).reshape(-1)
    if tol is None:
        rms = np.sqrt(np.mean(np.abs(ref) ** 2))
        tol = rel * np.maximum(np.abs(ref), rms)
    tol = np.asarray(tol, dtype=np.float64).reshape(-1)
    err = np.abs(dev - ref)
    worst = np.argmax(err / tol)
    assert np.all(err <= tol), '%s: worst err %.3g vs tol %.3g at %d' % (what, err[worst], tol[worst], worst)
    return float(err[worst] / tol[worst])


def eq_tolerance(det, rel, Nr, Nt):
    """Per-symbol error bound of the detected symbols z_k = G_k y_k (SISO: G_k = 1 / H_k, the one-tap equaliser)
    when the demodulated rx samples y_k and the channel H_k each carry a relative error `rel` (first order):
        |dz| <= ||G_k||_2 * rel * ( max(||y_k||_2, rms ||y||_2) + ||H_k||_F * ||z_k||_2 )
    (dz = G dy + dG y with dG = -G dH G, so dG y = -G dH z).  det: `detail` of oracle.links.ofdm_tdl_frame.
    Returns tol[n_sym*used*Nt] in the layout of the equalised symbols (symbol j*Nt + t)."""
    G, Hk = det['G'], det['Hk']                         # [K, Nt, Nr], [K, Nr, Nt]
    K = G.shape[0]
    y = np.moveaxis(det['Y'].reshape(Nr, K), 0, -1)    # [K, Nr]
    z = det['eq'].reshape(K, Nt)
    gn = np.linalg.norm(G, 2, axis=(1, 2))
    yn = np.linalg.norm(y, axis=1)
    yn = np.maximum(yn, np.sqrt(np.mean(yn ** 2)))
    hn = np.linalg.norm(Hk, 'fro', axis=(1, 2))
    zn = np.linalg.norm(z, axis=1)
    return np.repeat(rel * gn * (yn + hn * zn), Nt)


def decision_margin(modem, r):
    """Distance gap between the best and second-best constellation points (oracle side)."""
    if modem.kind == 'bpsk':
        return 2 * np.abs(np.real(r))
    d = np.abs(modem.symbols.reshape(-1, 1) - np.asarray(r).reshape(1, -1))
    d.sort(axis=0)
    return d[1] - d[0]


def assert_decisions(dev_idx, ref_idx, modem, ref_samples, exact, eps=2e-4, what=''):
    """exact: indices identical.  Otherwise (float32 arithmetic vs the float64 oracle) any mismatch
    must sit on a decision boundary: oracle margin (distance gap between the two nearest constellation
    points, which a sample error e can close only if it is < 2 e) below eps — a scalar or one value per
    symbol — and be rare."""
    dev_idx = np.asarray(dev_idx).reshape(-1).astype(np.int64)
    ref_idx = np.asarray(ref_idx).reshape(-1).astype(np.int64)
    bad = np.nonzero(dev_idx != ref_idx)[0]
    if exact:
        assert bad.size == 0, '%s: %d index mismatches (first at %s)' % (what, bad.size, bad[:5])
        return 0
    if bad.size:
        marg = decision_margin(modem, np.asarray(ref_samples).reshape(-1)[bad])
        lim = eps if np.isscalar(eps) else np.asarray(eps).reshape(-1)[bad]
        assert np.all(marg < lim), '%s: mismatch with margin %.3g (limit %.3g)' % (
            what, marg.max(), np.max(lim))
        assert bad.size <= max(2, 1e-3 * dev_idx.size), '%s: %d mismatches' % (what, bad.size)
    return bad.size


def host_draws(kind, seed, units, **kw):
    return getattr(OL, 'draws_' + kind)(seed=seed, units=units, **kw)


__all__ = ['philox']
