"""LinkSimulationRunner: the reference's Monte Carlo loop over a fused link op.

The simulators of the reference (apps/awgn_modulators/simulate_psk.py:51-115, apps/mimo/simulate_mimo.py:
68-142) are `SimulationRunner` subclasses whose `_run_simulation` simulates ONE small repetition in NumPy
and returns six `Result`s; `SimulationRunner.simulate()` (simulations/runner.py:1435-1539, 1670-1697) loops
repetitions and SNR points around it.  This class keeps that structure — the loop, `_keep_going`, partial
results, the six Results — and makes one repetition a *batch* of `units_per_rep` independent realizations
simulated by one fused-link call:

  * the call goes through the host-buffer C entry points in Monte Carlo mode (`b200phy_link_*_host` with no
    draw arrays): the parameters go in, 32 bytes of counters come back;
  * under torch.distributed every rank simulates its contiguous shard of the repetition's units
    (`distributed.shard`) and ONE all-reduce of the 4 counters makes all ranks build identical Results, so
    `_keep_going` and the partial-result files agree on every rank;
  * realization `u` of an SNR point is Philox unit `u`: repetition r covers units [r B, (r+1) B), so the
    counters after R repetitions are a pure function of (seed, SNR, R B) — independent of the batch size,
    of the number of GPUs and of where a resumed run picked up.
"""
import numpy as np

from .. import distributed
from .results import counters_to_results
from .runner import SimulationRunner

__all__ = ['LinkSimulationRunner']


class LinkSimulationRunner(SimulationRunner):
    """`link_call(noise_var, first_unit, n_units) -> int64[4]` (NumPy, host) simulates units
    [first_unit, first_unit + n_units) at the given noise variance, e.g.
    ``lambda nv, first, n: (link.set_noise_var(nv), link.run_host(n, first_unit=first))[1]``.

    Parameters `SNR` (unpacked, dB; noise variance = 1 / dB2Linear(SNR) as in simulate_psk.py:73) and
    `units_per_rep` are stored in `self.params` like the reference apps store `SNR` / `NSymbs`."""

    def __init__(self, link_call, units_per_rep, SNR, rep_max=1, max_bit_errors=None,
                 progressbar_message=None):
        super().__init__(read_command_line_args=False)
        self._link_call = link_call
        self.rep_max = rep_max
        self.max_bit_errors = max_bit_errors
        self.params.add('SNR', np.atleast_1d(np.asarray(SNR, dtype=float)))
        self.params.set_unpack_parameter('SNR')
        self.params.add('units_per_rep', int(units_per_rep))
        self.progressbar_message = progressbar_message
        if progressbar_message is None:
            self.update_progress_function_style = None
        self._cursor = 0

    # the unit cursor restarts with every SNR point and continues where a partial-result file stopped
    def _on_simulate_current_params_start(self, current_params):
        self._cursor = 0

    def _load_partial_results(self, current_params):
        loaded = super()._load_partial_results(current_params)
        if loaded is not None:
            self._cursor = int(loaded.current_rep) * int(current_params['units_per_rep'])
        return loaded

    def _run_simulation(self, current_parameters):
        B = int(current_parameters['units_per_rep'])
        noise_var = 1.0 / (10.0 ** (float(current_parameters['SNR']) / 10.0))
        first, n = distributed.shard(B, self._cursor)
        self._cursor += B
        counters = np.asarray(self._link_call(noise_var, first, n), dtype=np.int64)
        if distributed.world_size() > 1:
            import torch
            dev = torch.from_numpy(counters)
            if torch.distributed.get_backend() == 'nccl':
                dev = dev.cuda()
            counters = distributed.allreduce_counters(dev).cpu().numpy()
        return counters_to_results(counters)

    def _keep_going(self, current_params, current_sim_results, current_rep):
        if self.max_bit_errors is None:
            return True
        return current_sim_results['bit_errors'][-1].get_result() < self.max_bit_errors
