"""Mirror of pyphysim.channels for the single-link hot path (fading generators, TDL channel,
single-user channel wrappers)."""
from . import fading, fading_generators, singleuser  # noqa: F401
