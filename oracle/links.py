"""Per-realization link compositions (the five BASELINE.json configs) — NumPy.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each function composes the
restated reference stages exactly the way the reference's callers do:

  siso_flat   notebooks/Transmission_with_Rayleigh_and_AWGN_channels.ipynb cell 8
              (AWGN variant: apps/awgn_modulators/simulate_psk.py:51-115)
  alamouti / blast_flat   apps/mimo/simulate_mimo.py:68-142
  ofdm_tdl    notebooks/TDL_and_OFDM.ipynb cell 32 (SISO); the MIMO variant
              composes Blast.encode -> OFDM.modulate per antenna ->
              TdlMimoChannel.corrupt_data -> OFDM.demodulate per antenna ->
              per-subcarrier Blast.decode as SURVEY.md §8(d) C5 defines.

All random draws are passed in (see oracle/philox.py for how both sides derive
them); outputs are the demapped indices and the 4 counters
``[symbol_errors, bit_errors, num_symbols, num_bits]``.
"""
import math

import numpy as np

from . import fading, mimo, modulators as md, ofdm, philox


class Modem:
    """kind in {'qam', 'psk', 'bpsk'}; mirrors QAM(M) / PSK(M, offset) / BPSK()."""

    def __init__(self, kind, M=2, phase_offset=0.0):
        self.kind = kind
        if kind == 'qam':
            self.symbols = md.qam_constellation(M)
        elif kind == 'psk':
            self.symbols = md.psk_constellation(M, phase_offset)
        elif kind == 'bpsk':
            self.symbols = md.bpsk_constellation()
        else:
            raise ValueError(kind)
        self.M = self.symbols.size
        self.bits = md.level2bits(self.M)

    def modulate(self, idx):
        return md.bpsk_modulate(idx) if self.kind == 'bpsk' else md.modulate(self.symbols, idx)

    def demodulate(self, r):
        return md.bpsk_demodulate(r) if self.kind == 'bpsk' else md.demodulate(self.symbols, r)


def counters(idx, idx_hat, bits):
    """simulate_psk.py:83-86: symbol errors, bit errors, totals."""
    idx = np.asarray(idx).reshape(-1)
    idx_hat = np.asarray(idx_hat).reshape(-1)
    return np.array([int(np.sum(idx != idx_hat)), int(md.count_bit_errors(idx, idx_hat)),
                     idx.size, idx.size * bits], dtype=np.int64)


# ---- flat links -----------------------------------------------------------
def siso_flat(modem, idx, h, n, noise_var):
    """One symbol per realization: r = h x + sqrt(noise_var) n; r /= h (h None: AWGN)."""
    x = modem.modulate(idx)
    nn = math.sqrt(noise_var) * n
    if h is None:
        r = x + nn
    else:
        r = h * x + nn
        r = r / h
    return modem.demodulate(r), r


def alamouti(modem, idx, H, n, noise_var):
    """idx[U, S], H[U, Nr, 2], n[U, Nr, S] -> idx_hat[U, S], decoded[U, S]."""
    U, S = idx.shape
    dec = np.empty((U, S), dtype=complex)
    for u in range(U):
        x = mimo.alamouti_encode(modem.modulate(idx[u]))
        y = np.dot(H[u], x) + n[u] * np.sqrt(noise_var)
        dec[u] = mimo.alamouti_decode(y, H[u])
    return modem.demodulate(dec), dec


def blast_flat(modem, idx, H, n, noise_var, filter_noise_var=0.0):
    """idx[U, S*Nt], H[U, Nr, Nt], n[U, Nr, S].  filter_noise_var > 0 selects MMSE."""
    U = idx.shape[0]
    Nt = H.shape[2]
    dec = np.empty(idx.shape, dtype=complex)
    for u in range(U):
        x = mimo.blast_encode(modem.modulate(idx[u]), Nt)
        y = np.dot(H[u], x) + n[u] * np.sqrt(noise_var)
        dec[u] = mimo.blast_decode(y, H[u], filter_noise_var)
    return modem.demodulate(dec), dec


def precoded_flat(modem, scheme, idx, H, n, noise_var, filter_noise_var=0.0):
    """SVDMimo / GMDMimo / MRT over flat Rayleigh (apps/mimo/simulate_mimo.py:68-142): idx[U, S*layers] with
    symbol p = l*S + s on layer l (X = reshape(Nt, -1)), H[U, Nr, Nt], n[U, Nr, S].  The SVD is the
    gauge-fixed one (mimo.svd_canonical): the convention of the CUDA decomposition."""
    U = idx.shape[0]
    Nr, Nt = H.shape[1:]
    dec = np.empty(idx.shape, dtype=complex)
    for u in range(U):
        sym = modem.modulate(idx[u])
        if scheme == 'mrt':
            y = np.dot(H[u], mimo.mrt_encode(sym, H[u])) + n[u] * np.sqrt(noise_var)
            dec[u] = mimo.mrt_decode(y, H[u])
            continue
        Uc, Sc, Vc = mimo.svd_canonical(H[u])
        if scheme == 'svd':
            W = Vc / np.sqrt(Nt)
            G = np.diag(1.0 / Sc).dot(Uc.conj().T) * np.sqrt(Nt)
        else:
            Q, R, P = mimo.gmd(Uc, Sc, Vc.conj().T)
            W = P / np.sqrt(Nt)
            G = mimo.blast_receive_filter(Q.dot(R), filter_noise_var)
        y = np.dot(H[u], mimo.precoded_encode(sym, W)) + n[u] * np.sqrt(noise_var)
        dec[u] = G.dot(y).reshape(-1)
    return modem.demodulate(dec), dec


# ---- OFDM over a Jakes-TDL channel ---------------------------------------
class OfdmTdlConfig:
    def __init__(self, modem, fft, cp, used=None, n_sym=1, Nr=1, Nt=1,
                 profile=fading.COST259_TU, Ts=None, Fd=10.0, L=20, t0=None,
                 noise_var=0.01, filter_noise_var=None):
        self.modem = modem
        self.fft, self.cp, self.used = ofdm.check_parameters(fft, cp, used)
        self.n_sym, self.Nr, self.Nt = n_sym, Nr, Nt
        self.Ts = Ts if Ts is not None else 1.0 / (15e3 * fft)
        self.Fd, self.L = Fd, L
        # JakesSampleGenerator.__init__ emits one sample, so the first used
        # sample sits at t = Ts (fading_generators.py:351)
        self.t0 = self.Ts if t0 is None else t0
        self.tap_powers, self.delays = fading.discretize_profile(profile[0], profile[1], self.Ts)
        self.mem = int(self.delays[-1])
        self.noise_var = noise_var
        # Blast.set_noise_var value; None -> same as the channel noise variance,
        # 0 -> zero forcing (mimo.py:547-553, :597-605)
        self.filter_noise_var = noise_var if filter_noise_var is None else filter_noise_var
        self.N = n_sym * (self.fft + self.cp)
        self.n_data = Nt * n_sym * self.used
        self.mimo = not (Nr == 1 and Nt == 1)

    @property
    def phase_shape(self):
        ntaps = self.delays.size
        return (self.L, ntaps, self.Nr, self.Nt) if self.mimo else (self.L, ntaps)


def ofdm_tdl_frame(cfg, idx, phi, psi, noise, reference_equalizer=False, detail=False, per_subcarrier_loop=False):
    """One frame.  idx[n_data], phi/psi[cfg.phase_shape], noise[Nr, N+mem]
    (unit variance).  Returns idx_hat[n_data] (and intermediates if detail)."""
    m = cfg.modem
    s = m.modulate(idx)
    if cfg.mimo:
        layers = mimo.blast_encode(s, cfg.Nt)                               # [Nt, n_sym*used]
        tx = np.stack([ofdm.modulate(layers[t], cfg.fft, cfg.cp, cfg.used)
                       for t in range(cfg.Nt)])                            # [Nt, N]
    else:
        tx = ofdm.modulate(s, cfg.fft, cfg.cp, cfg.used)                    # [N]
    h, _ = fading.jakes_samples(phi, psi, cfg.Fd, cfg.Ts, cfg.t0, cfg.N)
    taps = fading.tdl_taps(h, cfg.tap_powers)
    rx = fading.tdl_corrupt(tx, taps, cfg.delays)
    rx = np.atleast_2d(rx) + math.sqrt(cfg.noise_var) * noise
    Y = np.stack([ofdm.demodulate(rx[r, :cfg.N], cfg.fft, cfg.cp, cfg.used)
                  for r in range(cfg.Nr)])                                  # [Nr, n_sym*used]
    if not cfg.mimo:
        if reference_equalizer:
            Hm = ofdm.mean_freq_response_reference(taps, cfg.delays, cfg.fft, cfg.n_sym)
        else:
            Hm = ofdm.mean_freq_response(taps, cfg.delays, cfg.fft, cfg.n_sym)
        eq = ofdm.onetap_equalize(Y[0], Hm, cfg.fft, cfg.used)
    else:
        Hm = ofdm.mean_freq_response(taps, cfg.delays, cfg.fft, cfg.n_sym)  # [n_sym, fft, Nr, Nt]
        bins = ofdm.used_subcarrier_indexes(cfg.fft, cfg.used)
        Yg = Y.reshape(cfg.Nr, cfg.n_sym, cfg.used)
        if per_subcarrier_loop:
            # exactly how a caller of the reference would do it: one Blast object state per subcarrier
            eq = np.empty(cfg.n_data, dtype=complex)
            for sy in range(cfg.n_sym):
                for q in range(cfg.used):
                    j = sy * cfg.used + q
                    eq[j * cfg.Nt:(j + 1) * cfg.Nt] = mimo.blast_decode(
                        Yg[:, sy, q].reshape(cfg.Nr, 1), Hm[sy, bins[q]], cfg.filter_noise_var)
        else:
            # same LAPACK calls on the stack of all subcarriers (mimo.blast_receive_filter_batched)
            Hk = Hm[:, bins]                                                 # [n_sym, used, Nr, Nt]
            G = mimo.blast_receive_filter_batched(Hk.reshape(-1, cfg.Nr, cfg.Nt), cfg.filter_noise_var)
            yk = np.moveaxis(Yg, 0, -1).reshape(-1, cfg.Nr, 1)               # [n_sym*used, Nr, 1]
            eq = (G @ yk).reshape(-1)                                        # symbol j*Nt + t
    idx_hat = m.demodulate(eq)
    if detail:
        det = dict(tx=tx, h=h, rx=rx, Y=Y, Hm=Hm, eq=eq)
        if cfg.mimo and not per_subcarrier_loop:
            det['G'] = G                                                     # [n_sym*used, Nt, Nr]
            det['Hk'] = Hk.reshape(-1, cfg.Nr, cfg.Nt)
        elif not cfg.mimo:
            Hk = Hm[:, ofdm.used_subcarrier_indexes(cfg.fft, cfg.used)].reshape(-1, 1, 1)
            det['G'], det['Hk'] = 1.0 / Hk, Hk                               # the one-tap "filter"
        return idx_hat, det
    return idx_hat


def ofdm_tdl(cfg, idx, phi, psi, noise, reference_equalizer=False):
    """Batch of frames: idx[U, n_data], phi/psi[U, *phase_shape], noise[U, Nr, N+mem]."""
    out = np.empty(idx.shape, dtype=np.int64)
    for u in range(idx.shape[0]):
        out[u] = ofdm_tdl_frame(cfg, idx[u], phi[u], psi[u], noise[u], reference_equalizer)
    return out


# ---- draws for a range of units (shared Philox stream) ---------------------
def draws_siso_flat(seed, units, bits, rayleigh=True, dtype=np.float64):
    idx = philox.data_indices(seed, units, 1, bits)[:, 0]
    h = philox.cnormal(seed, philox.STREAM_CHANNEL, units, 1, dtype)[:, 0] if rayleigh else None
    n = philox.cnormal(seed, philox.STREAM_NOISE, units, 1, dtype)[:, 0]
    return idx, h, n


def draws_flat_mimo(seed, units, bits, Nr, Nt, S, n_data, dtype=np.float64):
    idx = philox.data_indices(seed, units, n_data, bits)
    H = philox.cnormal(seed, philox.STREAM_CHANNEL, units, Nr * Nt, dtype).reshape(-1, Nr, Nt)
    n = philox.noise_rows(seed, units, Nr, S, dtype)
    return idx, H, n


def draws_ofdm_tdl(cfg, seed, units, dtype=np.float64):
    idx = philox.data_indices(seed, units, cfg.n_data, cfg.modem.bits)
    phi, psi = philox.jakes_phases(seed, units, cfg.phase_shape, dtype)
    noise = philox.noise_rows(seed, units, cfg.Nr, cfg.N + cfg.mem, dtype)
    return idx, phi, psi, noise
