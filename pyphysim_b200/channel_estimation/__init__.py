"""Mirror of pyphysim.channel_estimation."""
from . import estimators  # noqa: F401
from .estimators import (compute_ls_estimation, compute_mmse_estimation, compute_theoretical_ls_MSE,  # noqa: F401
                         compute_theoretical_mmse_MSE)
