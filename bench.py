#!/usr/bin/env python
"""bench.py — Monte Carlo realizations/s of the per-realization link hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one SNR point: one pass of the fused link kernel over a batch of synthetic frames per GPU, followed
(N>1) by one all-reduce of the 4 error counters.  Default workload = the one BASELINE.json's north_star quotes
the metric on: 64-QAM, 2x2 MIMO (Blast-MMSE), 1024-subcarrier OFDM over a Jakes/TDL (COST-259 TU) channel.
Prints ONE JSON line (rank 0).

  value       stream mode: the draws (data indices, Jakes phases, noise) are tensors resident in HBM, the kernel
              reads them and writes the demapped indices + counters (device-timed, CUDA events)
  roofline    algorithmic HBM bytes of stream mode / kernel time vs the measured copy peak; `traffic` and `issue`
              from the committed ncu capture of the SAME kernel instantiation (profiles/ncu_<workload>.json,
              checked against the name of the kernel this run launched)
  fused_rng   in-kernel Philox mode (no draw tensors at all), device-timed — the mode a Monte Carlo run uses
  e2e         the Monte Carlo mode END TO END through the public API: a `LinkSimulationRunner`
              (SimulationRunner subclass; Python owns the loop) runs K SNR points; every point is one
              `_run_simulation` -> `b200phy_link_*_host` (parameters in, kernel, 32 B of counters out over PCIe)
              -> counters all-reduce -> six `Result`s merged by the reference's loop.  Wall clock, max over ranks.
  e2e_stream  round-1's e2e: stream-mode work through the host-buffer entry point with pinned host draws
              (H2D of every draw inside the timed region: PCIe-bound)
  parity      float32 (timed) vs float64 (oracle-exact) kernels on the same device draws: counter drift
  f64         throughput of the float64 arithmetic (what bit-exact counts cost)
  configs     the other BASELINE.json configs (C2..C5) measured in the same run
  cpu_baseline / --impl reference: the NumPy oracle port of the reference path on the host cores
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: kind, M, fft, cp, used, n_sym, Nr, Nt, snr_dB, units/step/GPU
    'ofdm1024_qam64_mimo2x2_tdl': dict(kind='ofdm', M=64, fft=1024, cp=72, used=1024, n_sym=1, Nr=2, Nt=2,
                                       snr_dB=25.0, units=100000),
    'c3_ofdm1024_qam64_siso_tdl': dict(kind='ofdm', M=64, fft=1024, cp=72, used=1024, n_sym=1, Nr=1, Nt=1,
                                       snr_dB=20.0, units=100000),
    'c5_ofdm2048_qam256_mimo4x4_tdl': dict(kind='ofdm', M=256, fft=2048, cp=144, used=2048, n_sym=1, Nr=4,
                                           Nt=4, snr_dB=30.0, units=20000),
    # C2 is an SNR sweep (BASELINE.json: 0-30 dB): the e2e leg walks these points, the device-timed legs use snr_dB
    'c2_qam64_flat_rayleigh': dict(kind='siso_flat', M=64, snr_dB=15.0, units=100000000,
                                   snr_sweep_dB=[0.0, 5.0, 10.0, 15.0, 20.0, 25.0, 30.0]),
    'c4_qpsk_alamouti2x2': dict(kind='alamouti', M=4, Nr=2, S=2, snr_dB=10.0, units=20000000),
    # channel-dependent precoding, one 4x4 SVD (+ GMD) per realization, 16 symbol vectors
    'n3_qam16_svd4x4': dict(kind='precoded', scheme='svd', M=16, Nr=4, Nt=4, S=16, snr_dB=20.0, units=4000000),
    'n3_qam16_gmd4x4': dict(kind='precoded', scheme='gmd', M=16, Nr=4, Nt=4, S=16, snr_dB=20.0, units=4000000),
}
DEFAULT = 'ofdm1024_qam64_mimo2x2_tdl'
SIDE_CONFIGS = ['c2_qam64_flat_rayleigh', 'c3_ofdm1024_qam64_siso_tdl', 'c4_qpsk_alamouti2x2',
                'c5_ofdm2048_qam256_mimo4x4_tdl']
SEED = 0x5EEDB200
METRIC, UNIT = "monte_carlo_realizations_per_s", "realizations/s"


def dB2Linear(v):
    return 10.0 ** (v / 10.0)


def make_config(wname, w, world):
    """The `config` object: identical in the b200 arm and the reference arm (the driver compares them)."""
    return {"workload": wname, **{k: v for k, v in w.items() if k != 'kind'},
            "units_per_step_per_gpu": int(w['units']),
            "mode": "value: stream mode (draw tensors resident in HBM); e2e: Monte Carlo mode (in-kernel Philox) "
                    "through a SimulationRunner and the host-buffer C entry points",
            "l2": "stream-mode inputs are %.2f GB per step >> 126 MB L2: no flush needed" % (
                bytes_per_unit_of(w) * w['units'] / 1e9),
            "parallelism": "realizations sharded over %d GPU(s), one 32 B counter all-reduce per step" % world}


def bytes_per_unit_of(w):
    """Algorithmic HBM bytes per realization / frame of stream mode (idx 1 B, complex64 8 B,
    phase 4 B): idx + draws in, idx_hat out."""
    k = w['kind']
    if k == 'siso_flat':
        return 18
    if k == 'alamouti':
        return w['S'] + w['Nr'] * 2 * 8 + w['Nr'] * w['S'] * 8 + w['S']
    if k == 'precoded':
        return w['S'] * w['Nt'] + w['Nr'] * w['Nt'] * 8 + w['Nr'] * w['S'] * 8 + w['S'] * w['Nt']
    mem = {1024: 33, 2048: 66}.get(w['fft'])           # COST-259 TU at Ts = 1/(15 kHz fft): last tap delay
    if mem is None:
        mem = int(tu_profile(w['fft'])[0].tap_delays[-1])
    n_data = w['Nt'] * w['n_sym'] * w['used']
    return n_data + w['Nr'] * (w['n_sym'] * (w['fft'] + w['cp']) + mem) * 8 + 2 * 20 * 15 * w['Nr'] * w['Nt'] * 4 + n_data


# ------------------------------------------------------------------ clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        """One persistent `nvidia-smi -lms 50` child: a sample every 50 ms while the timed region runs."""
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        for line in self.proc.stdout:
            if self.stop_flag:
                break
            parts = [c.strip() for c in line.strip().split(',')]
            if len(parts) >= 7:
                self.rows.append(parts + [time.perf_counter()])

    def wait_first(self, timeout=3.0):
        """Block until nvidia-smi has delivered its first sample (it takes a few hundred ms to start)."""
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < timeout and self.is_alive():
            time.sleep(0.02)

    def begin(self):
        """Only samples taken after this call (i.e. under load) are summarised."""
        self.t_begin = time.perf_counter()

    def under_load(self):
        tb = getattr(self, 't_begin', 0.0)
        return [r for r in list(self.rows) if r[-1] >= tb]

    def stop(self):
        self.stop_flag = True
        proc = getattr(self, 'proc', None)
        if proc is not None:
            proc.terminate()
        self.join(timeout=2)

    def summary(self):
        rows = self.under_load()
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in rows if r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows)}


# ------------------------------------------------------------------ workload construction
def tu_profile(fft):
    from pyphysim_b200.channels.fading import COST259_TUx
    Ts = 1.0 / (15e3 * fft)
    return COST259_TUx.get_discretize_profile(Ts), Ts


def make_link(w, dtype='f32', tensor_cores=False):
    from pyphysim_b200 import links
    from pyphysim_b200.modulators import QAM
    prof, Ts = tu_profile(w['fft'])
    return links.OfdmTdlLink(QAM(w['M']), w['fft'], w['cp'], w['used'], num_ofdm_symbols=w['n_sym'],
                             Nr=w['Nr'], Nt=w['Nt'], tap_powers_linear=prof.tap_powers_linear,
                             tap_delays=prof.tap_delays, Fd=10.0, Ts=Ts, L=20,
                             noise_var=1.0 / dB2Linear(w['snr_dB']), dtype=dtype, seed=SEED,
                             use_tensor_cores=tensor_cores)


def measure_tensor_core_variant(w, R, steps, world, lib, base_value, base_fused):
    """The same workload with the per-subcarrier channel matrices on the tensor cores (tcgen05, 3xTF32, opt-in): its
    rates beside the default kernel's, its counters against the default kernel's on the same frames."""
    import torch
    link = make_link(w, tensor_cores=True)
    first = 0
    draws = link.draw(first, R)
    cnt = torch.zeros(4, dtype=torch.int64, device='cuda')
    stream = lambda: link.run(R, first_unit=first, draws=draws, counters=cnt)       # noqa: E731
    fused = lambda: link.run(R, first_unit=first, counters=cnt)                     # noqa: E731
    for _ in range(3):
        stream()
    ms_s = timed(stream, steps, world)
    kernel = lib.b200phy_last_kernel().decode()
    for _ in range(3):
        fused()
    ms_f = timed(fused, steps, world)
    c_tc = link.run(R, first_unit=first, draws=draws)
    c_cc = make_link(w).run(R, first_unit=first, draws=draws)
    del draws
    cap, _ = ncu_capture_for(DEFAULT + '_tcgen05', kernel)
    out = {"kernel": kernel, "value": R / (ms_s * 1e-3), "fused_rng": R / (ms_f * 1e-3), "unit": UNIT,
           "vs_default_stream": R / (ms_s * 1e-3) / base_value, "vs_default_fused": R / (ms_f * 1e-3) / base_fused,
           "symbol_count_drift_vs_default": int(c_tc[0]) - int(c_cc[0]), "bit_count_drift_vs_default": int(c_tc[1]) - int(c_cc[1]),
           "units": R, "adopted": False,
           "what": "H_k = sum_j gbar_j W^(k d_j) of a frame as one tcgen05 tile (M 128 x N 64 x K 32, 3xTF32, DFT operand "
                   "resident in TMEM, 12 MMAs per frame): fewer instructions, not faster (the kernel is bound by the FMA pipe "
                   "of the FIR); opt-in via OfdmTdlLink(use_tensor_cores=True) / params.reserved bit 1"}
    if cap:
        out["tensor_pipe_pct"] = cap.get('tensor_pipe_pct')
        out["warp_inst_per_unit"] = cap.get('warp_inst_per_unit')
    return out


def oracle_frame_runner(w):
    """The NumPy port of the reference path for workload w: returns f(units) -> counters."""
    from oracle import fading as ofading
    from oracle import links as OL
    if w['kind'] == 'ofdm':
        om = OL.Modem('qam', w['M'])
        cfg = OL.OfdmTdlConfig(om, w['fft'], w['cp'], w['used'], n_sym=w['n_sym'], Nr=w['Nr'], Nt=w['Nt'],
                               profile=ofading.COST259_TU, Fd=10.0, L=20,
                               noise_var=1.0 / dB2Linear(w['snr_dB']))

        def run(units):
            idx, phi, psi, noise = OL.draws_ofdm_tdl(cfg, SEED, units)
            # SISO: the reference's own per-sample-FFT equaliser (ofdm.py:541-548); MIMO: FFT(mean taps)
            hat = OL.ofdm_tdl(cfg, idx, phi, psi, noise, reference_equalizer=not cfg.mimo)
            return OL.counters(idx, hat, om.bits)
        return run
    if w['kind'] == 'siso_flat':
        om = OL.Modem('qam', w['M'])

        def run(units):
            tot = np.zeros(4, dtype=np.int64)
            for s in range(0, len(units), 1000):           # the reference apps use 1000-symbol calls
                idx, h, n = OL.draws_siso_flat(SEED, units[s:s + 1000], om.bits)
                hat, _ = OL.siso_flat(om, idx, h, n, 1.0 / dB2Linear(w['snr_dB']))
                tot += OL.counters(idx, hat, om.bits)
            return tot
        return run
    if w['kind'] == 'precoded':
        om = OL.Modem('qam', w['M'])

        def run(units):
            idx, H, n = OL.draws_flat_mimo(SEED, units, om.bits, w['Nr'], w['Nt'], w['S'], w['S'] * w['Nt'])
            hat, _ = OL.precoded_flat(om, w['scheme'], idx, H, n, 1.0 / dB2Linear(w['snr_dB']))
            return OL.counters(idx, hat, om.bits)
        return run
    om = OL.Modem('psk', 4, np.pi / 4)

    def run(units):
        idx, H, n = OL.draws_flat_mimo(SEED, units, 2, w['Nr'], 2, w['S'], w['S'])
        hat, _ = OL.alamouti(om, idx, H, n, 1.0 / dB2Linear(w['snr_dB']))
        return OL.counters(idx, hat, 2)
    return run


def _cpu_worker(args):
    wname, lo, hi = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    run = oracle_frame_runner(WORKLOADS[wname])
    t0 = time.perf_counter()
    c = run(np.arange(lo, hi))
    return time.perf_counter() - t0, c


def cpu_sample_size(w):
    """Units for the single-core cpu_baseline leg: about 10-20 s of NumPy work."""
    return {'ofdm': 256 if w.get('Nr', 1) * w.get('Nt', 1) <= 4 else 32, 'siso_flat': 4000000,
            'alamouti': 300000, 'precoded': 60000}[w['kind']]


def time_cpu(wname, cores, units_per_core):
    """Run the oracle port on `cores` processes, disjoint unit slices; returns (units/s, seconds)."""
    import multiprocessing as mp
    jobs = [(wname, i * units_per_core, (i + 1) * units_per_core) for i in range(cores)]
    # one BLAS/OpenMP thread per process: the workers are the parallelism (set before they start)
    for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMEXPR_NUM_THREADS'):
        os.environ[var] = '1'
    t0 = time.perf_counter()
    if cores == 1:
        try:
            from threadpoolctl import threadpool_limits
            with threadpool_limits(limits=1):          # this process's BLAS pool is already up: clamp it
                _cpu_worker(jobs[0])
        except ImportError:
            _cpu_worker(jobs[0])
    else:
        with mp.get_context('spawn').Pool(cores) as pool:
            # warm the workers (imports) before timing
            pool.map(_cpu_worker, [(wname, 0, 1)] * cores)
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs)
    dt = time.perf_counter() - t0
    return cores * units_per_core / dt, dt


def usable_cores():
    """Host threads this process may actually run on: affinity mask, capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    quota = None
    try:
        with open('/sys/fs/cgroup/cpu.max') as f:                      # cgroup v2: "<quota|max> <period>"
            q, per = f.read().split()
            if q != 'max':
                quota = float(q) / float(per)
    except (OSError, ValueError):
        try:
            with open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us') as f:     # cgroup v1
                q = float(f.read())
            with open('/sys/fs/cgroup/cpu/cpu.cfs_period_us') as f:
                per = float(f.read())
            if q > 0:
                quota = q / per
        except (OSError, ValueError):
            pass
    if quota:
        n = min(n, max(1, int(quota + 0.5)))
    return max(1, n)


def port_anchor(wname):
    """How the port relates to the imported reference (timed in the build container, tools/time_reference.py)."""
    try:
        d = json.load(open(os.path.join(ROOT, 'profiles', 'reference_vs_port_cpu_r02.json')))['configs'][wname]
        return {"port_over_unmodified_reference_1core": d['port_over_reference'],
                "source": "profiles/reference_vs_port_cpu_r02.json (build container, tools/time_reference.py)"}
    except Exception:
        return None


# ------------------------------------------------------------------ reference arm
def run_reference(args, wname, w, emit):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = usable_cores()
    per_core = max(1, cpu_sample_size(w) // 16)      # a few seconds per step on every core
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = time_cpu(wname, cores, per_core)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    sample = "%d units on each of %d processes per step (NumPy oracle port of the reference path)" % (per_core, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(wname, w, max(1, args.gpus)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "anchor": port_anchor(wname)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------ B200 arm: one workload
class Work:
    """Everything bench needs from one workload on this rank: the stream-mode step (draws resident in HBM), the
    fused-RNG step, the Monte Carlo host call for the runner, the stream-mode host call."""

    def __init__(self, wname, w, rank, lib):
        import torch
        from pyphysim_b200 import _lib, links
        from pyphysim_b200.modulators import QAM, QPSK
        self.wname, self.w, self.lib = wname, w, lib
        R = self.R = int(w['units'])
        first = self.first = rank * R                  # weak scaling: rank g owns units [g*R, (g+1)*R)
        self.counters = counters = torch.zeros(4, dtype=torch.int64, device='cuda')
        self.bytes_per_unit = bytes_per_unit_of(w)
        self.stream_host = None
        # bytes a Monte Carlo host call moves H2D: the constellation table (float pairs) + the parameter block
        self.mc_h2d_bytes = 2 * 4 * w['M'] + (C.sizeof(_lib.OfdmTdlParams) if w['kind'] == 'ofdm' else 64)
        kind = w['kind']
        if kind == 'ofdm':
            link = self.link = make_link(w)
            assert self.bytes_per_unit == link.bytes_per_frame(), (self.bytes_per_unit, link.bytes_per_frame())
            draws = self.draws = link.draw(first, R)   # resident in HBM before the timed region
            self.step = lambda: link.run(R, first_unit=first, draws=draws, counters=counters, want_idx=True)
            self.fused = lambda: link.run(R, first_unit=first, counters=counters)
            self.sym_per_unit = link.n_data

            def mc(noise_var, f0, n):
                link.set_noise_var(noise_var)
                return link.run_host(n, first_unit=f0)
            self.mc = mc
            Re = self.Re = min(R, 20000)
            self.make_host_draws = lambda: tuple(t[:Re].cpu().pin_memory() for t in draws)
            self.stream_host = lambda hd: link.run_host(Re, first_unit=first, draws=hd, want_idx=True)
            self.d2h_stream = Re * link.n_data + 32
            self.note = "OFDM/TDL frames are instruction-issue bound, not HBM bound (see `issue`, DESIGN.md, profiles/)"
        elif kind == 'siso_flat':
            mod = QAM(w['M'])
            nv = 1.0 / dB2Linear(w['snr_dB'])
            draws = self.draws = links.draw_siso_flat(mod, R, seed=SEED, first_unit=first)
            hat = torch.empty(R, dtype=torch.uint8, device='cuda')
            modem, self._keep = mod._native(_lib.F32)
            self.step = lambda: _lib.check(lib.b200phy_link_siso_flat(
                _lib.F32, modem, 1, nv, SEED, first, R, _lib.ptr(draws[0]), _lib.ptr(draws[1]), _lib.ptr(draws[2]),
                _lib.ptr(hat), None, _lib.ptr(counters), _lib.cur_stream()))
            self.fused = lambda: links.link_siso_flat(mod, nv, R, seed=SEED, first_unit=first, counters=counters)
            self.mc = lambda noise_var, f0, n: links.link_siso_flat_host(mod, noise_var, n, seed=SEED, first_unit=f0)
            Re = self.Re = min(R, 20000000)
            self.make_host_draws = lambda: tuple(t[:Re].cpu().pin_memory() for t in draws)
            self.stream_host = lambda hd: links.link_siso_flat_host(mod, nv, Re, seed=SEED, first_unit=first,
                                                                    draws=hd, want_idx=True)
            self.d2h_stream = Re + 32
            self.sym_per_unit = 1
            self.note = "HBM-bound elementwise link"
        elif kind == 'precoded':
            mod = QAM(w['M'])
            nv = 1.0 / dB2Linear(w['snr_dB'])
            S, Nr, Nt, sch = w['S'], w['Nr'], w['Nt'], w['scheme']
            draws = self.draws = links.draw_flat_mimo(mod, R, Nr=Nr, Nt=Nt, num_symbols=S, n_data=S * Nt, seed=SEED,
                                                      first_unit=first)
            kw = dict(scheme=sch, Nr=Nr, Nt=Nt, num_symbols=S)
            self.step = lambda: links.link_precoded(mod, nv, R, draws=draws, counters=counters, want_idx=True, **kw)
            self.fused = lambda: links.link_precoded(mod, nv, R, seed=SEED, first_unit=first, counters=counters, **kw)
            self.mc = lambda noise_var, f0, n: links.link_precoded_host(mod, noise_var, n, seed=SEED, first_unit=f0, **kw)
            Re = self.Re = min(R, 500000)
            self.make_host_draws = lambda: tuple(t[:Re].cpu().pin_memory() for t in draws)
            self.stream_host = lambda hd: links.link_precoded_host(mod, nv, Re, seed=SEED, first_unit=first, draws=hd,
                                                                   want_idx=True, **kw)
            self.d2h_stream = Re * S * Nt + 32
            self.sym_per_unit = S * Nt
            self.note = "bound by the per-realization double-precision Jacobi SVD (FP64 pipe), not HBM"
        else:
            mod = QPSK()
            nv = 1.0 / dB2Linear(w['snr_dB'])
            S, Nr = w['S'], w['Nr']
            draws = self.draws = links.draw_flat_mimo(mod, R, Nr=Nr, Nt=2, num_symbols=S, n_data=S, seed=SEED,
                                                      first_unit=first)
            hat = torch.empty((R, S), dtype=torch.uint8, device='cuda')
            modem, self._keep = mod._native(_lib.F32)
            self.step = lambda: _lib.check(lib.b200phy_link_alamouti(
                _lib.F32, modem, Nr, S, nv, SEED, first, R, _lib.ptr(draws[0]), _lib.ptr(draws[1]), _lib.ptr(draws[2]),
                _lib.ptr(hat), None, _lib.ptr(counters), _lib.cur_stream()))
            self.fused = lambda: links.link_alamouti(mod, nv, R, Nr=Nr, num_symbols=S, seed=SEED, first_unit=first,
                                                     counters=counters)
            self.mc = lambda noise_var, f0, n: links.link_alamouti_host(mod, noise_var, n, Nr=Nr, num_symbols=S,
                                                                        seed=SEED, first_unit=f0)
            Re = self.Re = min(R, 5000000)
            self.make_host_draws = lambda: tuple(t[:Re].cpu().pin_memory() for t in draws)
            self.stream_host = lambda hd: links.link_alamouti_host(mod, nv, Re, Nr=Nr, num_symbols=S, seed=SEED,
                                                                   first_unit=first, draws=hd, want_idx=True)
            self.d2h_stream = Re * S + 32
            self.sym_per_unit = S
            self.note = "HBM-bound elementwise link"

    def release(self):
        self.draws = None
        self.step = self.fused = self.stream_host = self.make_host_draws = None
        import torch
        torch.cuda.empty_cache()


def timed(fn, k, world):
    """CUDA-event time of k calls of fn on the current stream, max over ranks (ms per call)."""
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / k], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def wall_max(seconds, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([seconds], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def measure_e2e_runner(work, n_points, world, warm_points=2):
    """The Monte Carlo mode end to end: LinkSimulationRunner.simulate() over n_points SNR points, one batch of
    R units per rank and point.  Returns (seconds per SNR point, launches, final SimulationResults)."""
    import torch
    import torch.distributed as dist
    from pyphysim_b200.simulations import LinkSimulationRunner
    w = work.w
    sweep = w.get('snr_sweep_dB') or [w['snr_dB']]
    R, lib = work.R, work.lib

    def call(noise_var, f0, n):
        # LinkSimulationRunner shards [cursor, cursor + world*R) over the ranks: rank g gets R units
        return work.mc(noise_var, f0, n)

    def runner(points):
        snr = [sweep[i % len(sweep)] for i in range(points)]
        return LinkSimulationRunner(call, world * R, snr, rep_max=1)

    runner(warm_points).simulate()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    r = runner(n_points)
    l0 = lib.b200phy_launch_count()
    t0 = time.perf_counter()
    r.simulate()
    dt = time.perf_counter() - t0
    launches = int(lib.b200phy_launch_count() - l0)
    return wall_max(dt, world) / n_points, launches, r.results


def measure_e2e_stream(work, steps, world):
    """Round-1's e2e: stream-mode work through the host-buffer entry point, pinned host draws, H2D inside."""
    import torch
    import torch.distributed as dist
    hd = work.make_host_draws()
    h2d = sum(t.numel() * t.element_size() for t in hd if t is not None)
    for _ in range(2):
        work.stream_host(hd)
    if world > 1:
        dist.barrier()
    ke = max(3, min(steps, 5))
    t0 = time.perf_counter()
    for _ in range(ke):
        work.stream_host(hd)
    e2e_s = wall_max((time.perf_counter() - t0) / ke, world)
    # PCIe-bound: time a plain pinned H2D copy of the same size on this box beside it
    probe_h = torch.empty(min(int(h2d), 1 << 29), dtype=torch.uint8, pin_memory=True)
    probe_d = torch.empty(probe_h.numel(), dtype=torch.uint8, device='cuda')
    probe_d.copy_(probe_h, non_blocking=True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    p0.record()
    for _ in range(3):
        probe_d.copy_(probe_h, non_blocking=True)
    p1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * probe_h.numel() / (p0.elapsed_time(p1) * 1e-3) / 1e9
    del probe_h, probe_d, hd
    return {"value": world * work.Re / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(work.d2h_stream), "units_per_step_per_gpu": work.Re,
            "api": "b200phy_link_*_host (pinned host draws, stream mode)", "h2d_gbs_achieved": h2d / e2e_s / 1e9,
            "h2d_gbs_plain_memcpy": h2d_gbs, "pcie_frac": (h2d / e2e_s / 1e9) / h2d_gbs}


def ncu_capture_for(wname, kernel_launched):
    """profiles/ncu_<workload>.json (tools/ncu_to_json.py from an `ncu --set full` capture).  Refuses a capture of a
    different kernel instantiation than the one this run launched: a stale profile must not be quoted."""
    path = os.path.join(ROOT, 'profiles', 'ncu_%s.json' % wname)
    try:
        cap = json.load(open(path))
    except Exception:
        return None, "no capture at profiles/ncu_%s.json" % wname
    if cap.get('kernel') != kernel_launched:
        return None, "capture is of %s, this run launched %s: not quoted" % (cap.get('kernel'), kernel_launched)
    return cap, None


def measure_workload(work, steps, warm, world, quick, sampler=None):
    """Device-timed legs + the Monte Carlo e2e leg of one workload; returns the dict that becomes the JSON line
    (default workload) or an entry of `configs`."""
    import torch
    import torch.distributed as dist
    lib, w, R = work.lib, work.w, work.R
    counters = work.counters

    def full_step():
        counters.zero_()
        work.step()
        if world > 1:
            dist.all_reduce(counters)                  # the one collective of the path: 32 bytes

    for _ in range(warm):
        full_step()
    if sampler is not None:
        sampler.start()
        sampler.wait_first()
        full_step()                                    # GPU busy again before the sampled window opens
        sampler.begin()
    l0 = lib.b200phy_launch_count()
    ms_step = timed(full_step, steps, world)
    launches = int(lib.b200phy_launch_count() - l0)
    kernel_stream = lib.b200phy_last_kernel().decode()
    ms_kernel = timed(work.step, steps, world)         # the dominant kernel alone (no memset / all-reduce)
    for _ in range(warm):
        work.fused()
    # auxiliary leg: two timed blocks of K steps, the faster one is kept (a single block has shown 3 % run-to-run
    # scatter right after the stream legs; the e2e leg below is timed once, through the runner)
    ms_fused = min(timed(work.fused, steps, world), timed(work.fused, steps, world))
    kernel_fused = lib.b200phy_last_kernel().decode()
    if sampler is not None:
        # short workloads end before nvidia-smi's 50 ms period has produced enough samples: keep the SAME load
        # running (untimed) until there are a few, so the reported clocks are clocks under this load
        t_ext = time.perf_counter()
        while len(sampler.under_load()) < 4 and time.perf_counter() - t_ext < 2.0:
            work.step()
            torch.cuda.synchronize()
        sampler.stop()
    counters.zero_()
    work.step()
    final = counters.cpu().numpy().tolist()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak_gbs = float(peaks.get('hbm_gbs', 6650.0))
    bpu = work.bytes_per_unit
    achieved = bpu * R / (ms_kernel * 1e-3) / 1e9
    value = world * R / (ms_step * 1e-3)
    out = {
        "value": value, "unit": UNIT, "symbols_per_s": value * work.sym_per_unit, "ms_per_step": ms_step,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": None,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                     "bytes_per_unit": bpu, "algorithmic_bytes": bpu * R, "kernel_ms": ms_kernel,
                     "kernel": kernel_stream, "note": work.note},
        "fused_rng": {"value": world * R / (ms_fused * 1e-3), "unit": UNIT, "ms_per_step": ms_fused,
                      "kernel": kernel_fused,
                      "bound": "fp32 issue / SFU (no HBM traffic beyond 32 B of counters)"},
        "gpu_launches": launches, "counters": final,
    }
    cap, why = ncu_capture_for(work.wname, kernel_stream)
    if cap:
        wi = float(cap['warp_inst_per_unit'])
        sm_mhz = 1965.0
        if sampler is not None:
            sm_mhz = sampler.summary().get("sm_mhz") or 1965.0
        peak_wi = 148 * 4 * sm_mhz * 1e6
        out["roofline"]["traffic"] = cap['dram_bytes_per_unit'] * R
        out["roofline"]["traffic_note"] = (
            "dram__bytes_read.sum + dram__bytes_write.sum = %.0f B/unit in the ncu capture of %s (%d units/launch) x %d "
            "units of this launch" % (cap['dram_bytes_per_unit'], cap['kernel'], cap['units_per_launch'], R))
        out["issue"] = {"warp_inst_per_unit": wi, "achieved_warp_inst_per_s": value / world * wi,
                        "peak_warp_inst_per_s": peak_wi, "frac": value / world * wi / peak_wi,
                        "source": "ncu smsp__inst_executed.sum, profiles/ncu_%s.json" % work.wname}
    else:
        out["roofline"]["traffic_note"] = why
    capf, _ = ncu_capture_for(work.wname + '_fused', kernel_fused)
    if capf:
        sm_mhz = (sampler.summary().get("sm_mhz") if sampler is not None else None) or 1965.0
        peak_wi = 148 * 4 * sm_mhz * 1e6
        rate = R / (ms_fused * 1e-3)
        out["fused_rng"]["issue"] = {
            "warp_inst_per_unit": float(capf['warp_inst_per_unit']), "achieved_warp_inst_per_s": rate * capf['warp_inst_per_unit'],
            "peak_warp_inst_per_s": peak_wi, "frac": rate * capf['warp_inst_per_unit'] / peak_wi,
            "source": "ncu smsp__inst_executed.sum, profiles/ncu_%s_fused.json" % work.wname}

    if not quick:
        # ---- e2e: the Monte Carlo mode through the SimulationRunner and the host-buffer C entry points
        n_points = max(3, steps)
        s_point, mc_launches, res = measure_e2e_runner(work, n_points, world)
        e2e_value = world * R / s_point
        out["e2e"] = {
            "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(work.mc_h2d_bytes), "d2h_bytes_per_step": 32,
            "units_per_step_per_gpu": R, "snr_points": n_points, "ms_per_snr_point": s_point * 1e3,
            "kernel_ms_per_snr_point": ms_fused,
            "runner_overhead_ms_per_snr_point": s_point * 1e3 - ms_fused,
            "frac_of_device_fused_rate": e2e_value / (world * R / (ms_fused * 1e-3)),
            "gpu_launches": mc_launches,
            "api": "LinkSimulationRunner.simulate() -> _run_simulation -> b200phy_link_*_host (Monte Carlo mode: no "
                   "draw arrays; the C call copies its parameter block in and 4 int64 counters out) -> all-reduce "
                   "-> counters_to_results; one SNR point per step, R units per rank and point",
            "h2d_note": "the inputs of a Monte Carlo step are the link parameters: the constellation table (cudaMemcpy "
                        "H2D inside every host call) and the parameter block (kernel arguments); the draws are "
                        "generated in-kernel.  e2e_stream is the mode that ships every draw over PCIe",
            "ser_last_point": float(res.get_result_values_list('ser')[-1])}
        if work.stream_host is not None:
            out["e2e_stream"] = measure_e2e_stream(work, steps, world)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default=DEFAULT, choices=sorted(WORKLOADS))
    ap.add_argument('--units', type=int, default=0, help='realizations per step per GPU (0 = workload default)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-configs', action='store_true', help='skip the other BASELINE configs (C2..C5)')
    ap.add_argument('--quick', action='store_true',
                    help='kernel A/B runs: device-timed legs only (no e2e, parity, configs, cpu_baseline)')
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints there (e.g. NCCL's version banner,
    # seen on the multi-GPU boxes) is routed to stderr; emit() writes the line to the real stdout
    sys.stdout.flush()
    real_out = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_out, (json.dumps(line) + '\n').encode())

    wname, w = args.workload, dict(WORKLOADS[args.workload])
    if args.units:
        w['units'] = args.units
    if args.impl == 'reference':
        return run_reference(args, wname, w, emit)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from pyphysim_b200 import _lib
    lib = _lib.load()
    warm = max(3, args.warmup)

    work = Work(wname, w, rank, lib)
    sampler = ClockSampler(local)
    m = measure_workload(work, args.steps, warm, world, args.quick, sampler)
    line = {
        "metric": METRIC, "value": m.pop("value"), "unit": UNIT, "symbols_per_s": m.pop("symbols_per_s"),
        "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": m.pop("ms_per_step"),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(wname, w, world)}
    line.update(m)
    line["clocks"] = sampler.summary()

    if not args.quick and w['kind'] == 'ofdm':
        # ---- parity: the timed float32 kernel against the float64 kernel on the same device draws
        from pyphysim_b200.diagnostics import precision_drift
        n_par = min(100000, work.R)
        t0 = time.perf_counter()
        d = precision_drift(work.link, n_par, first_unit=work.first)
        d["seconds"] = time.perf_counter() - t0
        d["what"] = ("float32 kernel (timed above) vs float64 kernel (decision-exact against the NumPy oracle in "
                     "tests/) on the same %d frames of device draws; drift = f32 - f64 counters" % n_par)
        line["parity"] = d
        # ---- f64: what the bit-exact arithmetic costs (fused RNG, device-timed)
        l64 = work.link.with_dtype('f64')
        R64 = max(1000, work.R // 10)
        c64 = torch.zeros(4, dtype=torch.int64, device='cuda')
        f64 = lambda: l64.run(R64, first_unit=work.first, counters=c64)   # noqa: E731
        for _ in range(2):
            f64()
        ms64 = timed(f64, max(2, args.steps // 3), world)
        line["f64"] = {"value": world * R64 / (ms64 * 1e-3), "unit": UNIT, "units_per_step_per_gpu": R64,
                       "mode": "fused RNG, float64 arithmetic (complex128 like the reference)",
                       "kernel": lib.b200phy_last_kernel().decode(),
                       "slowdown_vs_f32_fused": (world * work.R / (line["fused_rng"]["ms_per_step"] * 1e-3)) /
                                                (world * R64 / (ms64 * 1e-3))}
    work.release()

    if not args.quick and wname == DEFAULT and world == 1:
        # ---- the tensor-core variant of the headline kernel (tcgen05 H_k, opt-in): measured beside the default
        line["tensor_core_variant"] = measure_tensor_core_variant(w, min(work.R, 50000), max(3, args.steps // 2), world, lib,
                                                                  line["value"], line["fused_rng"]["value"])

    if not (args.quick or args.no_configs) and wname == DEFAULT:
        cfgs = {}
        for name in SIDE_CONFIGS:
            ws = dict(WORKLOADS[name])
            wk = Work(name, ws, rank, lib)
            mm = measure_workload(wk, max(3, min(args.steps, 5)), 3, world, False)
            mm["config"] = make_config(name, ws, world)
            mm["dtype"] = "f32"
            wk.release()
            cfgs[name] = mm
        line["configs"] = cfgs

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    if not (args.no_cpu or args.quick) and world == 1:
        n1 = cpu_sample_size(w)
        v1, dt1 = time_cpu(wname, 1, n1)
        line["cpu_baseline"] = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "%d units of the same workload in %.1f s on 1 core (NumPy oracle port)" % (n1, dt1),
                                "anchor": port_anchor(wname)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
