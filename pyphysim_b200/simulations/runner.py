"""SimulationRunner: the Monte Carlo host loop, with the API of pyphysim.simulations.runner.

The loop stays Python ("Python host loop — stays Python in the build"); what changes is
what one `_run_simulation` call covers: with the fused link ops of `pyphysim_b200.links` one call
simulates a whole batch of realizations on the GPU and returns the 4 error counters, so the
per-repetition Python overhead of the reference amortises over the batch.  Semantics kept from the
reference (simulations/runner.py:1076-1945): cartesian unpacking of parameters, the
`_keep_going` early stop evaluated between repetitions, `SkipThisOne`, the `elapsed_time` and
`num_skipped_reps` results, partial results saved every 500 repetitions / 300 s and resumed on the
next `simulate()`, final results saved to pickle/JSON.
"""
import argparse
import os
import sys
from pathlib import Path
from time import time

from ..util.misc import pretty_time
from .parameters import SimulationParameters
from .results import Result, SimulationResults

__all__ = ['SimulationRunner', 'SkipThisOne', 'get_partial_results_filename', 'get_common_parser']


class SkipThisOne(Exception):
    """Raise inside `_run_simulation` to discard the current repetition; it is counted in the
    'num_skipped_reps' result (runner.py:151-185)."""

    def __init__(self, msg='Skip this one'):
        super().__init__(msg)
        self.msg = msg


def get_common_parser():
    """Command-line options shared by the reference's simulators (runner.py:44-98)."""
    if get_common_parser.parser is None:
        parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
        group = parser.add_argument_group('General')
        group.add_argument('-i', '--index', type=int, help="Simulate only this parameter variation")
        group.add_argument('-c', '--config', help="Configuration file")
        group.add_argument('-n', '--do-not-run', action='store_true', default=False,
                           help="Only load existing results instead of simulating")
        get_common_parser.parser = parser
    return get_common_parser.parser


get_common_parser.parser = None


def get_partial_results_filename(results_base_filename, current_params, partial_results_folder=None):
    """'<base>_unpack_<zero-padded index>.pickle' (runner.py:109-145)."""
    digits = len(str(current_params.get_num_unpacked_variations()))
    name = '{0}_unpack_{1}.pickle'.format(results_base_filename,
                                          str(current_params.unpack_index).zfill(digits))
    if partial_results_folder is not None:
        name = os.path.join(partial_results_folder, name)
    return name


class _TextProgress:
    """Minimal text progress bar (the reference's progressbar package is observability only)."""

    def __init__(self, total, message, out=None, width=50):
        self.total, self.message, self.out, self.width = max(int(total), 1), message, out or sys.stdout, width
        self._last = -1
        if message:
            self.out.write(message + '\n')

    def progress(self, count):
        filled = int(self.width * min(count, self.total) / self.total)
        if filled != self._last:
            self._last = filled
            self.out.write('\r[' + '*' * filled + ' ' * (self.width - filled) + ']')
            if count >= self.total:
                self.out.write('\n')
            self.out.flush()


class SimulationRunner:
    """Base class for Monte Carlo simulators: implement `_run_simulation` (and optionally
    `_keep_going`, the `_on_simulate_*` hooks), set `rep_max` and `params`, call `simulate()`."""

    def __init__(self, default_config_file=None, config_spec=None, read_command_line_args=True,
                 save_parsed_file=False):
        self.rep_max = 1
        self._runned_reps = []
        self._params = SimulationParameters()
        self._config_filename = default_config_file
        self._param_variation_index = None
        if default_config_file is not None:
            if read_command_line_args:
                args, _ = get_common_parser().parse_known_args()
                if args.config is not None:
                    self._config_filename = args.config
                self._param_variation_index = args.index
            self._params = SimulationParameters.load_from_config_file(self._config_filename, config_spec,
                                                                      save_parsed_file)
        # results / persistence
        self._results = SimulationResults()
        self._results_base_filename = None
        self._partial_files = []
        self.delete_partial_results_bool = False
        self.partial_results_folder = 'partial_results'
        self._last_save = time()
        # tracking
        self._elapsed_time = 0.0
        self._tic = 0.0
        self.progressbar_message = 'Progress'
        self.update_progress_function_style = 'text1'
        self.progress_output_type = 'screen'
        self._variation_pos = (0, 1)
        self._async_results = None

    def __repr__(self):
        extra = "" if self.results_filename is None else ", results_filename='%s'" % self.results_filename
        return "%s(rep_max=%s, num_params_variations=%s%s)" % (
            self.__class__.__name__, self.rep_max, self.params.get_num_unpacked_variations(), extra)

    # objects holding open files / device handles are not pickled with the runner
    # (cf. SimulationTracking.__getstate__, runner.py:276-283)
    def __getstate__(self):
        st = dict(self.__dict__)
        st['_async_results'] = None
        return st

    # ---- properties ----------------------------------------------------------------------------
    params = property(lambda self: self._params)
    results = property(lambda self: self._results)
    runned_reps = property(lambda self: self._runned_reps)

    @property
    def elapsed_time(self):
        return pretty_time(self._elapsed_time)

    def set_results_filename(self, filename=None):
        self._results_base_filename = filename

    @property
    def results_base_filename(self):
        if self._results_base_filename is None:
            return None
        return self._results.get_filename_with_replaced_params(self._results_base_filename)

    @property
    def results_filename(self):
        base = self.results_base_filename
        if base is None:
            return None
        return base if os.path.splitext(base)[-1] == '.pickle' else base + '.pickle'

    def clear(self):
        self._runned_reps = []
        self._results = SimulationResults()
        self._elapsed_time = 0.0

    # ---- to be implemented / overridden by subclasses ------------------------------------------
    def _run_simulation(self, current_parameters):
        raise NotImplementedError("'_run_simulation' must be implemented in a subclass of SimulationRunner")

    def _keep_going(self, current_params, current_sim_results, current_rep):
        return True

    def _on_simulate_start(self):
        pass

    def _on_simulate_finish(self):
        pass

    def _on_simulate_current_params_start(self, current_params):
        pass

    def _on_simulate_current_params_finish(self, current_params, current_params_sim_results):
        pass

    # ---- partial results -----------------------------------------------------------------------
    def _partial_name(self, current_params):
        return get_partial_results_filename(self.results_base_filename, current_params,
                                            self.partial_results_folder)

    def _save_partial_results(self, current_rep, current_params, current_sim_results):
        if self._results_base_filename is None:
            return None
        current_sim_results.set_parameters(current_params)
        current_sim_results.current_rep = current_rep
        name = self._partial_name(current_params)
        if self.partial_results_folder is not None:
            os.makedirs(self.partial_results_folder, exist_ok=True)
        saved = Path(current_sim_results.save_to_file(name)).absolute()
        if saved not in self._partial_files:
            self._partial_files.append(saved)
        return saved.name

    def _save_partial_results_maybe(self, current_rep, current_params, current_sim_results):
        now = time()
        if now - self._last_save > 300 or current_rep % 500 == 0:
            self._save_partial_results(current_rep, current_params, current_sim_results)
            self._last_save = now

    def _load_partial_results(self, current_params):
        """runner.py:1019-1069: resume, or ValueError if the stored parameters differ."""
        if self._results_base_filename is None:
            return None
        name = self._partial_name(current_params)
        try:
            loaded = SimulationResults.load_from_file(name)
        except (IOError, OSError):
            return None
        if not current_params == loaded.params:
            raise ValueError("Partial results loaded from file does not match current parameters. \n"
                             "file: '{0}'\nDelete that file first to simulate with new "
                             "configuration.".format(name))
        return loaded

    # ---- the loop ------------------------------------------------------------------------------
    def _timed_run(self, current_params):
        tic = time()
        res = self._run_simulation(current_params)
        res.add_result(Result.create('elapsed_time', Result.SUMTYPE, time() - tic))
        return res

    def _progress_func(self, current_params):
        style = self.update_progress_function_style
        if style is None or self.progressbar_message is None:
            return lambda value: None
        try:
            msg = self.progressbar_message.format(**current_params.parameters)
        except (KeyError, IndexError):
            msg = self.progressbar_message
        pos, n = self._variation_pos
        if n > 1:
            msg = "%s (%d/%d)" % (msg, pos + 1, n)
        out = sys.stdout
        if self.progress_output_type == 'file':
            out = open('{0}_progress.txt'.format(self.results_base_filename or 'simulation'), 'a')
        return _TextProgress(self.rep_max, msg, out).progress

    def _simulate_for_current_params_common(self, current_params, update_progress_func=lambda value: None):
        """runner.py:1435-1539."""
        self._on_simulate_current_params_start(current_params)
        current_sim_results = self._load_partial_results(current_params)
        if current_sim_results is None:
            current_sim_results = self._timed_run(current_params)
            current_rep = 1
        else:
            current_rep = current_sim_results.current_rep
        current_sim_results.add_new_result("num_skipped_reps", Result.SUMTYPE, 0)
        while self._keep_going(current_params, current_sim_results, current_rep) and current_rep < self.rep_max:
            try:
                current_sim_results.merge_all_results(self._timed_run(current_params))
                current_rep += 1
                update_progress_func(current_rep)
            except SkipThisOne:
                current_sim_results['num_skipped_reps'][-1].update(1)
            self._save_partial_results_maybe(current_rep, current_params, current_sim_results)
        update_progress_func(self.rep_max)
        self._on_simulate_current_params_finish(current_params, current_sim_results)
        partial_name = self._save_partial_results(current_rep, current_params, current_sim_results)
        return current_rep, current_sim_results, partial_name

    def _simulate_common_setup(self):
        self.clear()
        self.params.add('rep_max', self.rep_max) if 'rep_max' not in self.params.parameters else \
            self.params.__setitem__('rep_max', self.rep_max)
        self._tic = time()
        self._results.rep_max = self.rep_max
        self._results.set_parameters(self.params)
        self._on_simulate_start()

    def simulate_common_cleaning(self):
        self._on_simulate_finish()
        self._elapsed_time = time() - self._tic
        self._results.runned_reps = self._runned_reps
        self._results.elapsed_time = self.elapsed_time
        if self._results_base_filename is not None:
            self._results.save_to_file(self._results_base_filename)
            if self.delete_partial_results_bool:
                for name in self._partial_files:
                    try:
                        os.remove(name)
                    except OSError:
                        pass
                self._partial_files = []

    def simulate(self, param_variation_index=None):
        """Run every parameter variation serially (or only `param_variation_index`, whose partial
        result file is then meant to be combined later — runner.py:1638-1736)."""
        if param_variation_index is None:
            param_variation_index = self._param_variation_index
        self._simulate_common_setup()
        variations = self.params.get_unpacked_params_list()
        if param_variation_index is not None:
            if self.results_base_filename is None:
                raise RuntimeError('The results filename must be set before calling the "simulate" method.')
            idx = int(param_variation_index)
            if 0 <= idx < len(variations):
                self._variation_pos = (idx, len(variations))
                rep, _, _ = self._simulate_for_current_params_common(variations[idx],
                                                                     self._progress_func(variations[idx]))
                self._runned_reps = rep
            return
        for pos, current_params in enumerate(variations):
            self._variation_pos = (pos, len(variations))
            rep, sim_results, _ = self._simulate_for_current_params_common(
                current_params, self._progress_func(current_params))
            self._runned_reps.append(rep)
            self._results.append_all_results(sim_results)
        self.simulate_common_cleaning()

    # ---- task-farm variant ---------------------------------------------------------------------
    @staticmethod
    def _simulate_for_current_params_parallel(obj, current_params, update_progress_func=None):
        return obj._simulate_for_current_params_common(current_params, update_progress_func or (lambda v: None))

    def simulate_in_parallel(self, view=None, wait=True):
        """One parameter variation per worker through an ipyparallel-style `view.map`
        (runner.py:1774-1855).  On a multi-GPU box prefer realization sharding
        (pyphysim_b200.distributed), which parallelises *inside* every variation."""
        if view is None:
            raise RuntimeError("simulate_in_parallel needs a load-balanced view (ipyparallel is not "
                               "a dependency of pyphysim_b200)")
        self._simulate_common_setup()
        variations = self.params.get_unpacked_params_list()
        self._async_results = view.map(SimulationRunner._simulate_for_current_params_parallel,
                                       [self] * len(variations), variations, block=False)
        if wait:
            self.wait_parallel_simulation()

    def wait_parallel_simulation(self):
        """runner.py:1857-1886."""
        if self._async_results is None:
            raise RuntimeError("wait_parallel_simulation method should only be called after "
                               "the simulate_in_parallel method.")
        results = self._async_results.get() if hasattr(self._async_results, 'get') else list(self._async_results)
        for reps, sim_results, filename in results:
            self._runned_reps.append(reps)
            self._results.append_all_results(sim_results)
            if filename is not None:
                self._partial_files.append(Path(self.partial_results_folder or '.', filename).absolute())
        self._async_results = None
        self.simulate_common_cleaning()
