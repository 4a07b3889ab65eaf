"""Mirror of pyphysim.channels: fading generators, TDL channel, single-user channel wrappers and the
multi-link grids (MuChannel / MuMimoChannel)."""
from . import fading, fading_generators, multiuser, singleuser  # noqa: F401
