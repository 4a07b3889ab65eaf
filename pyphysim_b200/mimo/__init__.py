"""Mirror of pyphysim.mimo: Blast ZF/MMSE, Alamouti, MRT, MRC, SVDMimo, GMDMimo and the SINR helpers."""
from .mimo import (GMDMimo, MRC, MRT, Alamouti, Blast, MimoBase, MisoBase, SVDMimo,  # noqa: F401
                   calc_post_processing_linear_SINRs, calc_post_processing_SINRs)

__all__ = ['MimoBase', 'MisoBase', 'Blast', 'Alamouti', 'MRT', 'MRC', 'SVDMimo', 'GMDMimo',
           'calc_post_processing_SINRs', 'calc_post_processing_linear_SINRs']
