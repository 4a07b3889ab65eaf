#!/usr/bin/env python
"""Run under torchrun on N GPUs: every rank simulates its shard of one SNR point with the fused links,
the counters are all-reduced over NCCL, and rank 0 checks them against the unsharded run on one GPU
(must be bit-identical: the Philox stream is keyed by the global realization index).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_check.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pyphysim_b200 import distributed as D   # noqa: E402
from pyphysim_b200 import links              # noqa: E402
from pyphysim_b200.channels.fading import COST259_TUx   # noqa: E402
from pyphysim_b200.modulators import QAM, QPSK           # noqa: E402


def main():
    world = D.init()
    rank = D.rank()
    ok = True
    # OFDM 2x2 headline shape
    Ts = 1.0 / (15e3 * 1024)
    prof = COST259_TUx.get_discretize_profile(Ts)
    link = links.OfdmTdlLink(QAM(64), 1024, 72, 1024, Nr=2, Nt=2, tap_powers_linear=prof.tap_powers_linear,
                             tap_delays=prof.tap_delays, Fd=10.0, Ts=Ts, L=20, noise_var=0.003, seed=77)
    total = 20001
    first, count = D.shard(total, first_unit=500)
    c = torch.zeros(4, dtype=torch.int64, device='cuda')
    link.run(count, first_unit=first, counters=c)
    D.allreduce_counters(c)
    c2 = torch.zeros(4, dtype=torch.int64, device='cuda')
    f2, n2 = D.shard(10 ** 7 + 3)
    links.link_alamouti(QPSK(), 0.1, n2, first_unit=f2, counters=c2)
    D.allreduce_counters(c2)
    c3 = torch.zeros(4, dtype=torch.int64, device='cuda')
    f3, n3 = D.shard(10 ** 7 + 1)
    links.link_siso_flat(QAM(64), 0.03, n3, first_unit=f3, counters=c3)
    D.allreduce_counters(c3)
    # SISO OFDM in float32 (frame-pair kernel) with an odd total: the shards are odd-sized and start at odd units,
    # so frames change lane / pair partner / masked-tail status between the sharded and the single-GPU run
    siso = links.OfdmTdlLink(QAM(64), 1024, 72, 1024, tap_powers_linear=prof.tap_powers_linear,
                             tap_delays=prof.tap_delays, Fd=10.0, Ts=Ts, L=20, noise_var=0.01, seed=78)
    c4 = torch.zeros(4, dtype=torch.int64, device='cuda')
    f4, n4 = D.shard(20003, first_unit=1)
    siso.run(n4, first_unit=f4, counters=c4)
    D.allreduce_counters(c4)
    # C5 shape: 4x4 MMSE over 2048-subcarrier OFDM / TDL, 256-QAM (the 512-thread antenna-pair kernel)
    Ts5 = 1.0 / (15e3 * 2048)
    prof5 = COST259_TUx.get_discretize_profile(Ts5)
    c5link = links.OfdmTdlLink(QAM(256), 2048, 144, 2048, Nr=4, Nt=4, tap_powers_linear=prof5.tap_powers_linear,
                               tap_delays=prof5.tap_delays, Fd=10.0, Ts=Ts5, L=20, noise_var=1e-3, seed=79)
    c5 = torch.zeros(4, dtype=torch.int64, device='cuda')
    f5, n5 = D.shard(4003, first_unit=9)
    c5link.run(n5, first_unit=f5, counters=c5)
    D.allreduce_counters(c5)
    # the Monte Carlo runner: every rank drives the same LinkSimulationRunner, results must agree with 1 GPU
    from pyphysim_b200.simulations import LinkSimulationRunner

    def mc(nv, f0, n):
        link.set_noise_var(nv)
        return link.run_host(n, first_unit=f0)
    runner = LinkSimulationRunner(mc, 6001, [20.0, 25.0], rep_max=2)
    runner.simulate()
    r_se = runner.results.get_result_values_list('symbol_errors')
    torch.cuda.synchronize()
    if rank == 0:
        link.set_noise_var(0.003)                  # the runner above walked the SNR points on this link object
        ref = link.run(total, first_unit=500)
        ref2 = links.link_alamouti(QPSK(), 0.1, 10 ** 7 + 3)
        ref3 = links.link_siso_flat(QAM(64), 0.03, 10 ** 7 + 1)
        ref4 = siso.run(20003, first_unit=1)
        ref5 = c5link.run(4003, first_unit=9)
        ref_se = []
        for snr in (20.0, 25.0):
            link.set_noise_var(10 ** (-snr / 10))
            ref_se.append(int(link.run(2 * 6001)[0]))
        same = r_se == ref_se
        ok &= same
        print('runner    world=%d sharded=%s single=%s %s' % (world, r_se, ref_se, 'OK' if same else 'MISMATCH'))
        for name, a, b in (('ofdm2x2', c, ref), ('alamouti', c2, ref2), ('siso', c3, ref3), ('ofdm_siso_f32_odd', c4, ref4),
                           ('c5_4x4', c5, ref5)):
            same = np.array_equal(a.cpu().numpy(), b)
            ok &= same
            print('%-9s world=%d sharded=%s single=%s %s' % (name, world, a.cpu().numpy().tolist(), list(b),
                                                            'OK' if same else 'MISMATCH'))
    if D.is_initialized():
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
