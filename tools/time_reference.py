#!/usr/bin/env python
"""Anchor of the CPU arm: the UNMODIFIED reference (imported from /root/reference) timed next to the NumPy oracle
port that bench.py's `cpu_baseline` / `--impl reference` legs run on the GPU box (the reference cannot travel).

    python tools/time_reference.py [--frames-c3 N] [--frames-2x2 N] > profiles/reference_vs_port_cpu_r02.json

Build container only.  Both sides run on ONE core (OMP/BLAS threads = 1), same process, same shapes:
  * C3       QAM64 / OFDM(1024, 72) / COST-259 TU / Jakes(10 Hz, L = 20), one OFDM symbol per frame, composed as
             notebooks/TDL_and_OFDM.ipynb cell 32 and driven by the reference's SimulationRunner.simulate()
             (one frame per `_run_simulation`, simulations/runner.py:1491-1517);
  * 2x2      the north-star workload: Blast.encode -> OFDM.modulate per tx antenna -> TdlMimoChannel.corrupt_data ->
             OFDM.demodulate per rx antenna -> Blast.set_channel_matrix / decode per subcarrier with
             H_k = mean over the symbol of TdlImpulseResponse.get_freq_response (the per-sample FFT the reference's
             own equaliser uses, modulators/ofdm.py:541-548).
The reference draws from NumPy's global MT19937, the port from the shared Philox stream: the work per frame is the
same, the random numbers are not (timing only; parity is pinned by tests/golden/)."""
import argparse
import json
import os
import sys
import tempfile
import time

for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMEXPR_NUM_THREADS'):
    os.environ[_v] = '1'

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
_stub = tempfile.mkdtemp()
open(os.path.join(_stub, 'validate.py'), 'w').write(
    'class VdtTypeError(Exception): pass\nclass VdtValueTooSmallError(Exception): pass\n'
    'class VdtValueTooBigError(Exception): pass\ndef is_float(v, *a, **k): return float(v)\n'
    'def is_integer(v, *a, **k): return int(v)\n')
sys.path.insert(0, _stub)
sys.path.insert(0, '/root/reference')

from pyphysim.channels import fading, fading_generators  # noqa: E402
from pyphysim.mimo import mimo as rmimo  # noqa: E402
from pyphysim.modulators import fundamental, ofdm as rofdm  # noqa: E402
from pyphysim.simulations import Result, SimulationResults, SimulationRunner  # noqa: E402
from pyphysim.util import misc  # noqa: E402
from pyphysim.util.conversion import dB2Linear  # noqa: E402

sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload table + the port's frame runner)


def results(idx, hat, bits):
    se = int(np.sum(idx != hat))
    be = int(misc.count_bit_errors(idx, hat))
    r = SimulationResults()
    r.add_new_result('symbol_errors', Result.SUMTYPE, se)
    r.add_new_result('num_symbols', Result.SUMTYPE, idx.size)
    r.add_new_result('bit_errors', Result.SUMTYPE, be)
    r.add_new_result('num_bits', Result.SUMTYPE, idx.size * bits)
    r.add_new_result('ber', Result.RATIOTYPE, be, idx.size * bits)
    r.add_new_result('ser', Result.RATIOTYPE, se, idx.size)
    return r


class RefOfdmTdl(SimulationRunner):
    """One frame per repetition, objects built per frame like notebook TDL_and_OFDM cell 32."""

    def __init__(self, w, frames):
        super().__init__(read_command_line_args=False)
        self.w = w
        self.rep_max = frames
        self.params.add('SNR', w['snr_dB'])
        self.update_progress_function_style = None
        self.qam = fundamental.QAM(w['M'])
        self.ofdm = rofdm.OFDM(w['fft'], w['cp'], w['used'])
        self.Ts = 1.0 / (15e3 * w['fft'])
        self.prof = fading.COST259_TUx.get_discretize_profile(self.Ts)
        self.bits = int(np.log2(w['M']))

    def _run_simulation(self, current_parameters):
        w, o, qam = self.w, self.ofdm, self.qam
        Nr, Nt, used, fft = w['Nr'], w['Nt'], w['used'], w['fft']
        nv = 1.0 / dB2Linear(current_parameters['SNR'])
        idx = np.random.randint(0, w['M'], Nt * used)
        if Nr == 1 and Nt == 1:
            jakes = fading_generators.JakesSampleGenerator(10.0, self.Ts, 20)
            ch = fading.TdlChannel(jakes, self.prof)
            tx = o.modulate(qam.modulate(idx))
            rx = ch.corrupt_data(tx)
            rx = rx + np.sqrt(nv) * misc.randn_c(rx.size)
            Y = o.demodulate(rx[0:tx.size].copy())
            eq = rofdm.OfdmOneTapEqualizer(o).equalize_data(Y, ch.get_last_impulse_response())
        else:
            jakes = fading_generators.JakesSampleGenerator(10.0, self.Ts, 20, shape=(Nr, Nt))
            ch = fading.TdlMimoChannel(jakes, self.prof)
            blast = rmimo.Blast()
            blast.set_channel_matrix(np.eye(Nt, dtype=complex))
            layers = blast.encode(qam.modulate(idx))
            tx = np.stack([o.modulate(layers[t]) for t in range(Nt)])
            rx = ch.corrupt_data(tx)
            rx = rx + np.sqrt(nv) * misc.randn_c(*rx.shape)
            Y = np.stack([o.demodulate(rx[r, 0:tx.shape[1]].copy()) for r in range(Nr)])
            Hf = ch.get_last_impulse_response().get_freq_response(fft)       # [fft, Nr, Nt, N]: FFT per sample
            Hf = Hf.mean(axis=-1)
            bins = o.get_used_subcarrier_indexes()
            eq = np.empty(idx.size, dtype=complex)
            blast.set_noise_var(nv)
            for q in range(used):
                blast.set_channel_matrix(Hf[bins[q]])
                eq[Nt * q:Nt * q + Nt] = blast.decode(Y[:, q].reshape(Nr, 1))
        return results(idx, qam.demodulate(eq), self.bits)


def time_ref(w, frames):
    r = RefOfdmTdl(w, frames)
    t0 = time.perf_counter()
    r.simulate()
    dt = time.perf_counter() - t0
    ser = r.results.get_result_values_list('ser')[0]
    return frames / dt, dt, ser


def time_port(wname, frames):
    run = bench.oracle_frame_runner(bench.WORKLOADS[wname])
    t0 = time.perf_counter()
    c = run(np.arange(frames))
    dt = time.perf_counter() - t0
    return frames / dt, dt, c[0] / c[2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames-c3', type=int, default=24)
    ap.add_argument('--frames-2x2', type=int, default=6)
    a = ap.parse_args()
    import platform
    out = {"what": "frames/s on ONE core of the build container: unmodified reference (SimulationRunner.simulate(), "
                   "one frame per repetition) vs the NumPy oracle port bench.py times on the GPU box",
           "cpu": platform.processor() or open('/proc/cpuinfo').read().split('model name')[1].split('\n')[0].strip(': \t'),
           "numpy": np.__version__, "threads": 1, "configs": {}}
    for wname, frames in (('c3_ofdm1024_qam64_siso_tdl', a.frames_c3), ('ofdm1024_qam64_mimo2x2_tdl', a.frames_2x2)):
        w = bench.WORKLOADS[wname]
        time_ref(w, 1)                                   # warm numba / imports
        rv, rdt, rser = time_ref(w, frames)
        pv, pdt, pser = time_port(wname, max(frames, 16))
        out["configs"][wname] = {
            "reference_frames_per_s": rv, "reference_seconds": rdt, "reference_frames": frames, "reference_ser": rser,
            "port_frames_per_s": pv, "port_seconds": pdt, "port_frames": max(frames, 16), "port_ser": pser,
            "port_over_reference": pv / rv,
            "port_equaliser": "reference's per-sample FFT (ofdm.py:541-548)" if w['Nr'] * w['Nt'] == 1
            else "FFT(mean taps) — mean_n FFT(h_n) == FFT(mean_n h_n), pinned at 1e-12 in tests/test_oracle_golden.py"}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
