"""Mirror of pyphysim.util for the hot path (conversion, misc)."""
from . import conversion, misc  # noqa: F401
