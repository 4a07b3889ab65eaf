"""Mirror of pyphysim.mimo for the hot path (Blast ZF/MMSE, Alamouti)."""
from .mimo import Alamouti, Blast, MimoBase  # noqa: F401

__all__ = ['MimoBase', 'Blast', 'Alamouti']
