"""Mirror of pyphysim.simulations: the Monte Carlo host loop and its containers."""
from .parameters import SimulationParameters, combine_simulation_parameters  # noqa: F401
from .results import Result, SimulationResults, counters_to_results  # noqa: F401
from .runner import SimulationRunner, SkipThisOne, get_partial_results_filename  # noqa: F401
from .linkrunner import LinkSimulationRunner  # noqa: F401

__all__ = ['SimulationRunner', 'LinkSimulationRunner', 'SimulationParameters', 'SimulationResults', 'Result', 'SkipThisOne',
           'combine_simulation_parameters', 'counters_to_results', 'get_partial_results_filename']
