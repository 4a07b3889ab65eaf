#!/usr/bin/env python
"""bench.py — Monte Carlo realizations/s of the per-realization link hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one SNR point: one pass of the fused link kernel over a batch of synthetic frames per
GPU, followed (N>1) by one all-reduce of the 4 error counters.  Default workload = the one
BASELINE.json's north_star quotes the metric on: 64-QAM, 2x2 MIMO (Blast-MMSE), 1024-subcarrier OFDM
over a Jakes/TDL (COST-259 TU) channel.  Prints ONE JSON line (rank 0).

  value     stream mode: the draws (data indices, Jakes phases, noise) are tensors resident in HBM,
            the kernel reads them and writes the demapped indices + counters (device-timed)
  e2e       the same stream-mode work through the host-buffer C entry point
            (b200phy_link_ofdm_tdl_host): pinned host draws, H2D, kernel, D2H inside the timed region
  fused_rng in-kernel Philox mode (no draw tensors at all) — reported beside, with its own bound
  roofline  algorithmic HBM bytes of stream mode / kernel time vs the measured copy peak
  cpu_baseline / --impl reference: the NumPy oracle port of the reference path on the host cores
"""
import argparse
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
    'c2_qam64_flat_rayleigh': dict(kind='siso_flat', M=64, snr_dB=15.0, units=100000000),
    'c4_qpsk_alamouti2x2': dict(kind='alamouti', M=4, Nr=2, S=2, snr_dB=10.0, units=20000000),
    # channel-dependent precoding, one 4x4 SVD (+ GMD) per realization, 16 symbol vectors
    'n3_qam16_svd4x4': dict(kind='precoded', scheme='svd', M=16, Nr=4, Nt=4, S=16, snr_dB=20.0, units=4000000),
    'n3_qam16_gmd4x4': dict(kind='precoded', scheme='gmd', M=16, Nr=4, Nt=4, S=16, snr_dB=20.0, units=4000000),
}
DEFAULT = 'ofdm1024_qam64_mimo2x2_tdl'
SEED = 0x5EEDB200


def dB2Linear(v):
    return 10.0 ** (v / 10.0)


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


def make_link(w, dtype='f32'):
    from pyphysim_b200 import links
    from pyphysim_b200.modulators import QAM
    prof, Ts = tu_profile(w['fft'])
    return links.OfdmTdlLink(QAM(w['M']), w['fft'], w['cp'], w['used'], num_ofdm_symbols=w['n_sym'],
                             Nr=w['Nr'], Nt=w['Nt'], tap_powers_linear=prof.tap_powers_linear,
                             tap_delays=prof.tap_delays, Fd=10.0, Ts=Ts, L=20,
                             noise_var=1.0 / dB2Linear(w['snr_dB']), dtype=dtype, seed=SEED)


def oracle_frame_runner(w):
    """The NumPy port of the reference path for workload w: returns f(units) -> counters."""
    from oracle import fading as ofading
    from oracle import links as OL
    from oracle import modulators as md
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
                res = [_cpu_worker(jobs[0])]
        except ImportError:
            res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context('spawn').Pool(cores) as pool:
            # warm the workers (imports) before timing
            pool.map(_cpu_worker, [(wname, 0, 1)] * cores)
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, jobs)
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
    line = {"impl": "reference", "metric": "monte_carlo_realizations_per_s", "value": value,
            "unit": "realizations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, **{k: v for k, v in w.items() if k not in ('kind', 'units')}},
            "cpu_baseline": {"value": value, "unit": "realizations/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "realizations/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------ B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default=DEFAULT, choices=sorted(WORKLOADS))
    ap.add_argument('--units', type=int, default=0, help='realizations per step per GPU (0 = workload default)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--quick', action='store_true', help='kernel A/B runs: skip the e2e and cpu_baseline legs')
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
    from pyphysim_b200 import _lib, links
    from pyphysim_b200.modulators import QAM, QPSK
    lib = _lib.load()
    R = int(w['units'])
    first = rank * R                                   # weak scaling: rank g owns units [g*R, (g+1)*R)
    warm = max(3, args.warmup)

    counters = torch.zeros(4, dtype=torch.int64, device='cuda')
    if w['kind'] == 'ofdm':
        link = make_link(w)
        bytes_per_unit = link.bytes_per_frame()
        draws = link.draw(first, R)                    # resident in HBM before the timed region
        step = lambda: link.run(R, first_unit=first, draws=draws, counters=counters, want_idx=True)  # noqa: E731
        fused = lambda: link.run(R, first_unit=first, counters=counters)                      # noqa: E731
        Re = min(R, 20000)
        host_draws = tuple(t[:Re].cpu().pin_memory() for t in draws)
        e2e_call = lambda: link.run_host(Re, first_unit=first, draws=host_draws, want_idx=True)  # noqa: E731
        h2d = sum(t.numel() * t.element_size() for t in host_draws)
        d2h = Re * link.n_data + 32
        sym_per_unit = link.n_data
    elif w['kind'] == 'siso_flat':
        mod = QAM(w['M'])
        nv = 1.0 / dB2Linear(w['snr_dB'])
        bytes_per_unit = 18
        draws = links.draw_siso_flat(mod, R, seed=SEED, first_unit=first)
        hat = torch.empty(R, dtype=torch.uint8, device='cuda')
        step = lambda: _siso_step(lib, links, mod, nv, R, first, draws, hat, counters)         # noqa: E731
        fused = lambda: links.link_siso_flat(mod, nv, R, seed=SEED, first_unit=first, counters=counters)  # noqa: E731
        Re = min(R, 20000000)
        host_draws = tuple(t[:Re].cpu().pin_memory() for t in draws)
        e2e_call = lambda: links.link_siso_flat_host(mod, nv, Re, seed=SEED, first_unit=first, draws=host_draws, want_idx=True)  # noqa: E731
        h2d = sum(t.numel() * t.element_size() for t in host_draws)
        d2h = Re + 32
        sym_per_unit = 1
    elif w['kind'] == 'precoded':
        mod = QAM(w['M'])
        nv = 1.0 / dB2Linear(w['snr_dB'])
        S, Nr, Nt, sch = w['S'], w['Nr'], w['Nt'], w['scheme']
        bytes_per_unit = S * Nt + Nr * Nt * 8 + Nr * S * 8 + S * Nt          # idx + H + noise in, idx_hat out
        draws = links.draw_flat_mimo(mod, R, Nr=Nr, Nt=Nt, num_symbols=S, n_data=S * Nt, seed=SEED, first_unit=first)
        kw = dict(scheme=sch, Nr=Nr, Nt=Nt, num_symbols=S)
        step = lambda: links.link_precoded(mod, nv, R, draws=draws, counters=counters, want_idx=True, **kw)  # noqa: E731
        fused = lambda: links.link_precoded(mod, nv, R, seed=SEED, first_unit=first, counters=counters, **kw)  # noqa: E731
        e2e_call, Re, h2d, d2h = None, 0, 0, 0
        sym_per_unit = S * Nt
    else:
        mod = QPSK()
        nv = 1.0 / dB2Linear(w['snr_dB'])
        S, Nr = w['S'], w['Nr']
        bytes_per_unit = S + Nr * 2 * 8 + Nr * S * 8 + S
        draws = links.draw_flat_mimo(mod, R, Nr=Nr, Nt=2, num_symbols=S, n_data=S, seed=SEED, first_unit=first)
        step = lambda: links.link_alamouti(mod, nv, R, Nr=Nr, num_symbols=S, draws=draws, counters=counters, want_idx=True)  # noqa: E731
        fused = lambda: links.link_alamouti(mod, nv, R, Nr=Nr, num_symbols=S, seed=SEED, first_unit=first, counters=counters)  # noqa: E731
        e2e_call, Re, h2d, d2h = None, 0, 0, 0
        sym_per_unit = S

    def full_step():
        counters.zero_()
        step()
        if world > 1:
            dist.all_reduce(counters)                  # the one collective of the path: 32 bytes

    def timed(fn, k):
        """CUDA-event time of k calls of fn on the current stream, max over ranks (ms per call)."""
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

    for _ in range(warm):
        full_step()
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first()
    full_step()                                        # GPU busy again before the sampled window opens
    sampler.begin()
    l0 = lib.b200phy_launch_count()
    ms_step = timed(full_step, args.steps)
    launches = int(lib.b200phy_launch_count() - l0)
    ms_kernel = timed(step, args.steps)                # the dominant kernel alone (no memset/all-reduce)
    for _ in range(warm):
        fused()
    ms_fused = timed(fused, args.steps)
    # short workloads end before nvidia-smi's 50 ms period has produced enough samples: keep the SAME
    # load running (untimed) until there are a few, so the reported clocks are clocks under this load
    t_ext = time.perf_counter()
    while len(sampler.under_load()) < 4 and time.perf_counter() - t_ext < 2.0:
        step()
        torch.cuda.synchronize()
    sampler.stop()
    final = counters.cpu().numpy().tolist()

    # e2e: host buffers -> H2D -> kernel -> D2H through the C entry point (wall clock incl. sync)
    e2e = None
    if e2e_call is not None and not args.quick:
        for _ in range(2):
            e2e_call()
        if world > 1:
            dist.barrier()
        ke = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(ke):
            e2e_call()
        dt = torch.tensor([(time.perf_counter() - t0) / ke], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_s = float(dt.item())
        # the e2e path is PCIe-bound: time a plain pinned H2D copy of the same size on this box beside it
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
        del probe_h, probe_d
        e2e = {"value": world * Re / e2e_s, "unit": "realizations/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "units_per_step_per_gpu": Re, "api": "b200phy_link_*_host (pinned host draws, stream mode)",
               "h2d_gbs_achieved": h2d / e2e_s / 1e9, "h2d_gbs_plain_memcpy": h2d_gbs,
               "pcie_frac": (h2d / e2e_s / 1e9) / h2d_gbs}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak_gbs = float(peaks.get('hbm_gbs', 6650.0))
    achieved = bytes_per_unit * R / (ms_kernel * 1e-3) / 1e9
    value = world * R / (ms_step * 1e-3)
    line = {
        "metric": "monte_carlo_realizations_per_s", "value": value, "unit": "realizations/s",
        "symbols_per_s": value * sym_per_unit, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wname, **{k: v for k, v in w.items() if k not in ('kind',)},
                   "units_per_step_per_gpu": R, "mode": "stream (draw tensors resident in HBM)",
                   "l2": "inputs %.2f GB per step >> 126 MB L2, no flush needed" % (bytes_per_unit * R / 1e9),
                   "parallelism": "realizations sharded over %d GPU(s), one 32 B counter all-reduce per step" % world},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": None,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                     "bytes_per_unit": bytes_per_unit, "kernel_ms": ms_kernel,
                     "note": ("OFDM/TDL frames are instruction-issue bound, not HBM bound (see `issue`, DESIGN.md, "
                              "profiles/)") if w['kind'] == 'ofdm' else
                             ("bound by the per-realization double-precision Jacobi SVD (FP64 pipe), not HBM"
                              if w['kind'] == 'precoded' else "HBM-bound elementwise link")},
        "fused_rng": {"value": world * R / (ms_fused * 1e-3), "unit": "realizations/s", "ms_per_step": ms_fused,
                      "bound": "fp32 issue / SFU (no HBM traffic beyond 32 B of counters)"},
        "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(), "counters": final,
    }
    if wname == DEFAULT:
        # the real bound of this kernel is instruction issue: instructions and DRAM bytes per frame come from
        # the committed ncu --set full capture of the same kernel (profiles/headline_kernel_ncu.json), scaled
        # to this launch's unit count; the rate is live
        try:
            cap = json.load(open(os.path.join(ROOT, 'profiles', 'headline_kernel_ncu.json')))
        except Exception:
            cap = None
        if cap:
            wi = float(cap['warp_inst_per_unit'])
            sm_mhz = line["clocks"].get("sm_mhz") or 1965.0
            peak_wi = 148 * 4 * sm_mhz * 1e6
            line["roofline"]["traffic"] = cap['dram_bytes_per_unit'] * R
            line["roofline"]["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum = %.0f B/frame in the ncu capture "
                                                "(%d frames/launch) x %d frames of this launch; algorithmic %d B/frame"
                                                % (cap['dram_bytes_per_unit'], cap['units_per_launch'], R, bytes_per_unit))
            line["roofline"]["algorithmic_bytes"] = bytes_per_unit * R
            line["issue"] = {"warp_inst_per_unit": wi, "achieved_warp_inst_per_s": value / world * wi,
                             "peak_warp_inst_per_s": peak_wi, "frac": value / world * wi / peak_wi,
                             "source": "ncu smsp__inst_executed.sum, profiles/headline_kernel_ncu.json"}
    if not (args.no_cpu or args.quick):
        n1 = cpu_sample_size(w)
        v1, dt1 = time_cpu(wname, 1, n1)
        line["cpu_baseline"] = {"value": v1, "unit": "realizations/s", "cores": 1, "kind": "port",
                                "sample": "%d units of the same workload in %.1f s on 1 core (NumPy oracle port)" % (n1, dt1)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def _siso_step(lib, links, mod, nv, R, first, draws, hat, counters):
    import ctypes as C
    from pyphysim_b200 import _lib
    modem, keep = mod._native(_lib.F32)
    _lib.check(lib.b200phy_link_siso_flat(_lib.F32, modem, 1, nv, SEED, first, R, _lib.ptr(draws[0]),
                                          _lib.ptr(draws[1]), _lib.ptr(draws[2]), _lib.ptr(hat), None,
                                          _lib.ptr(counters), _lib.cur_stream()))


if __name__ == '__main__':
    main()
