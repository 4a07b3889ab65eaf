"""MuChannel / MuMimoChannel — NumPy restatement of the link grid.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference: ``pyphysim/channels/multiuser.py:42-586``: a
num_rx x num_tx grid of independent SuChannel links with one power delay profile; receiver rx gets the sum
over transmitters of signal[tx] through link (rx, tx), each scaled by sqrt(pathloss[rx, tx])
(``channels/singleuser.py:130-151``).
"""
import numpy as np

from . import fading


class LinkGrid:
    """State of the grid: per link the Jakes phases and the generator clock (every generator starts at
    t = Ts: its constructor emits one sample, fading_generators.py:351)."""

    def __init__(self, phi, psi, Fd, Ts, tap_powers, delays, pathloss=None):
        # phi, psi: [num_rx, num_tx, L, taps(, Nr, Nt)]
        self.phi, self.psi = np.asarray(phi), np.asarray(psi)
        self.num_rx, self.num_tx = self.phi.shape[:2]
        self.Fd, self.Ts = Fd, Ts
        self.tap_powers, self.delays = tap_powers, delays
        self.pathloss = np.ones((self.num_rx, self.num_tx)) if pathloss is None else np.asarray(pathloss)
        self.t = np.full((self.num_rx, self.num_tx), Ts)
        self.last_taps = {}

    def _link(self, rx, tx, x, freq, switched):
        phi, psi = self.phi[rx, tx], self.psi[rx, tx]
        if freq is None:
            n = x.shape[-1]
            h, self.t[rx, tx] = fading.jakes_samples(phi, psi, self.Fd, self.Ts, self.t[rx, tx], n)
        else:
            fft_size, car = freq
            block = fft_size if car is None else len(car)
            h, self.t[rx, tx] = fading.jakes_block_samples(phi, psi, self.Fd, self.Ts, self.t[rx, tx],
                                                           x.shape[-1] // block, fft_size)
        taps = fading.tdl_taps(h, self.tap_powers)
        self.last_taps[(rx, tx)] = np.sqrt(self.pathloss[rx, tx]) * taps
        if switched and taps.ndim == 4:
            taps = np.swapaxes(taps, 1, 2)             # the reverse link sees H^T (fading.py:1098-1100)
        y = fading.tdl_corrupt(x, taps, self.delays) if freq is None else \
            fading.tdl_corrupt_freq(x, taps, self.delays, freq[0], freq[1])
        return np.sqrt(self.pathloss[rx, tx]) * y

    def corrupt(self, signal, freq=None, switched=False):
        """signal[tx] (2-D array or object array) -> list of per-receiver outputs.  switched: transmitters
        and receivers exchange roles (multiuser.py:276-279), link (rx, tx) serves pair (tx, rx)."""
        n_out, n_in = (self.num_tx, self.num_rx) if switched else (self.num_rx, self.num_tx)
        out = []
        for o in range(n_out):
            acc = 0
            for i in range(n_in):
                rx, tx = (i, o) if switched else (o, i)
                acc = acc + self._link(rx, tx, np.asarray(signal[i]), freq, switched)
            out.append(acc)
        return out
