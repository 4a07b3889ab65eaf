"""Mirror of pyphysim.reference_signals: Zadoff-Chu root sequences, SRS / DMRS user
sequences and the CAZAC-based channel estimators, generated and evaluated on the GPU."""
from . import channel_estimation, dmrs, root_sequence, srs, zadoffchu  # noqa: F401
