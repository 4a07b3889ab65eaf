"""The N>1 path on CPU: two gloo ranks shard the realizations of one SNR point, all-reduce the four
counters, and must reproduce the single-process counters exactly (the Philox stream is keyed by the
global realization index).  The per-rank arithmetic here is the NumPy oracle — this test covers the
host logic (sharding, collective, runner agreement), the GPU tests cover the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from pyphysim_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _link_counters(first, count):
    from oracle import links as OL
    om = OL.Modem('qam', 16)
    if count == 0:
        return np.zeros(4, dtype=np.int64)
    idx, h, n = OL.draws_siso_flat(99, np.arange(first, first + count), om.bits)
    hat, _ = OL.siso_flat(om, idx, h, n, 0.05)
    return OL.counters(idx, hat, om.bits)


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    assert D.init('gloo') == world and D.rank() == rank and D.world_size() == world
    first, count = D.shard(total, first_unit=1000)
    counters = torch.from_numpy(_link_counters(first, count))
    D.allreduce_counters(counters)

    # a runner whose _run_simulation covers a sharded batch: every rank must take the same
    # _keep_going decisions because every rank sees the all-reduced counters
    from pyphysim_b200.simulations import SimulationRunner, counters_to_results

    class Sharded(SimulationRunner):
        def __init__(self):
            super().__init__(read_command_line_args=False)
            self.rep_max = 6
            self.update_progress_function_style = None
            self.batch = 500
            self.rep = 0

        def _run_simulation(self, current_parameters):
            f, c = D.shard(self.batch, first_unit=self.rep * self.batch)
            self.rep += 1
            t = torch.from_numpy(_link_counters(f, c))
            return counters_to_results(D.allreduce_counters(t).numpy())

        def _keep_going(self, current_params, current_sim_results, current_rep):
            return current_sim_results['symbol_errors'][-1].get_result() < 250

    r = Sharded()
    r.simulate()
    np.save(os.path.join(out_dir, 'rank%d.npy' % rank),
            np.r_[counters.numpy(), r.runned_reps[0], r.results['symbol_errors'][0].get_result()])
    torch.distributed.destroy_process_group()


def test_shard_covers_every_unit_once():
    for total in (0, 1, 7, 1000, 12345):
        for world in (1, 2, 3, 8):
            pieces = [D.shard(total, 10, g, world) for g in range(world)]
            assert pieces[0][0] == 10 and sum(c for _, c in pieces) == total
            for (f0, c0), (f1, _) in zip(pieces, pieces[1:]):
                assert f0 + c0 == f1
    assert D.world_size() == 1 and D.rank() == 0
    t = torch.arange(4)
    assert D.allreduce_counters(t) is t                  # no-op without a process group


@pytest.mark.timeout(300)
def test_two_rank_counters_match_single_process(tmp_path):
    total, world = 3001, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    ref = _link_counters(1000, total)
    r0, r1 = (np.load(str(tmp_path / ('rank%d.npy' % g))) for g in range(world))
    assert np.array_equal(r0, r1)                        # identical results on every rank
    assert np.array_equal(r0[:4], ref) and ref[0] > 0    # == unsharded counters, bit exact
    assert r0[4] < 6 and r0[5] >= 250                    # early stop taken consistently
